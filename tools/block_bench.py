"""Individual block time-steps against the shared adaptive block step (tupan's ahermite) on the
same Plummer sphere: steps, particle updates, pair evaluations and wall time to a common t_end.

    python tools/block_bench.py [n] [order] [t_end]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tupan_b200 import backend, ics  # noqa: E402
from tupan_b200.block import BlockHermite  # noqa: E402
from tupan_b200.integrator import Integrator  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    order = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    t_end = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0 ** -4
    eta = 1.0 / 64
    lib = backend.require_gpu("float64")

    b = BlockHermite(eta, ics.make_plummer(n, seed=1), order=order, dt_max=t_end)
    ke0, pe0 = b.energies()
    torch.cuda.synchronize()
    l0 = lib.tupan_cuda_launch_count()
    t0 = time.perf_counter()
    b.evolve(t_end)
    torch.cuda.synchronize()
    wall_b = time.perf_counter() - t0
    ke1, pe1 = b.energies()
    evals = 2 * (2 if order == 4 else 3) + 1       # per block step: pec x (acc_jerk [+ snap_crackle]) + tstep
    print("block  N=%d order %d t_end=%g: %d block steps, %d particle steps (%.2f%% of steps x N), "
          "%.3e pairs, %.3f s, %d launches, energy error %.2e"
          % (n, order, t_end, b.block_steps, b.particle_steps, 100.0 * b.particle_steps / (b.block_steps * n),
             float(b.particle_steps) * n * (evals / 2.0), wall_b, lib.tupan_cuda_launch_count() - l0,
             ((ke1 + pe1) - (ke0 + pe0)) / (ke0 + pe0)), flush=True)

    it = Integrator(eta, 0.0, ics.make_plummer(n, seed=1), method="ahermite%d" % order)
    ke0, pe0 = it.energies()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    steps = it.evolve(t_end, check_every=16)
    torch.cuda.synchronize()
    wall_s = time.perf_counter() - t0
    ke1, pe1 = it.energies()
    per = 3 * (1 if order == 4 else 2) + 1
    print("shared N=%d order %d t_end=%g: %d steps (every particle each step), %.3e pairs, %.3f s, "
          "energy error %.2e" % (n, order, t_end, steps, float(steps) * n * n * per, wall_s,
                                 ((ke1 + pe1) - (ke0 + pe0)) / (ke0 + pe0)), flush=True)
    print("wall-time ratio shared / block: %.2f" % (wall_s / wall_b))


if __name__ == "__main__":
    main()
