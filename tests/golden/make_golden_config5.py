"""Golden runs for BASELINE.json configs[4]: binary-rich Plummer N = 16384 (8192 binaries,
tupan_b200.ics.make_binary_rich(seed=1), eps2 = 0), fp64, on the reference's C backend:

    asakura       eta = 1/64,  t_end = 2^-9   (Sakura/Kepler pairwise propagation, adaptive step)
    sakura        eta = 2^-11, t_end = 2^-10  (two shared steps)
    sia21s.kdk    eta = 2^-11, t_end = 2^-10, pn_order = 7, clight = 128  (post-Newtonian kicks, two
                  shared steps: the adaptive variant takes the binaries' 1.5e-5 steps, i.e. an hour of
                  host time per 128 steps at 632 flop per pair)

The reference's own single-threaded Python driver needs minutes per force evaluation at this size
(2.7e8 pairs per kernel call), so -- as for configs[1] (make_golden_config2.py) -- the runs are
driven by oracle/integrators.py (pinned bit-for-bit to the reference's integrators,
tests/test_oracle_integrators.py) calling the UNMODIFIED reference C backend (oracle/_ref) on
contiguous i-slices from a thread pool: per particle the same arithmetic as the reference's loop.

    python tests/golden/make_golden_config5.py [threads]

Stores energies before/after, step count, final clock and the final state of the 2 x 128 members of
128 evenly spaced binaries; the inputs are regenerated in the test from make_binary_rich(seed=1)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import integrators as oi  # noqa: E402
from tupan_b200 import ics  # noqa: E402

N = 16384
VEC = ("rx", "ry", "rz", "vx", "vy", "vz")
CASES = [("asakura", 1.0 / 64, 2.0 ** -9, None), ("sakura", 2.0 ** -11, 2.0 ** -10, None),
         ("sia21s.kdk", 2.0 ** -11, 2.0 ** -10, (7, 128.0))]


def main():
    threads = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    kind = "ref" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libtupan_ref_fp64.so")) else "oracle"
    src = ics.make_binary_rich(N, seed=1)
    ins = {k: getattr(src, k).copy() for k in ("mass", "eps2") + VEC}
    pairs = np.linspace(0, N // 2 - 1, 128).astype(np.int64)
    idx = np.sort(np.concatenate([2 * pairs, 2 * pairs + 1]))
    flat = {"idx": idx}
    for method, eta, t_end, pn in CASES:
        t0 = time.time()
        b0 = oi.Bodies(ins, "float64", kind, threads)
        ke0, pe0 = b0.kinetic_energy, b0.potential_energy
        ps, steps = oi.evolve(ins, "float64", method, eta, t_end, kind=kind, threads=threads, pn=pn)
        ke1, pe1 = ps.kinetic_energy, ps.potential_energy
        for k in VEC:
            flat["%s/out/%s" % (method, k)] = ps.a[k][idx]
        flat["%s/meta" % method] = np.array([eta, t_end, steps, float(ps.clock[0]), ke0, pe0, ke1, pe1,
                                             pn[0] if pn else 0, pn[1] if pn else 0.0])
        print(method, "steps", steps, "t", float(ps.clock[0]), "eerr", ((ke1 + pe1) - (ke0 + pe0)) / (-pe1),
              "%.0fs" % (time.time() - t0), kind, flush=True)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "config5_binary_rich_n16384.npz"),
                        **flat)


if __name__ == "__main__":
    main()
