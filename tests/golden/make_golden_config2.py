"""Golden run for BASELINE.json configs[1]: Plummer N = 65536, shared-time-step leapfrog
(sia21s.dkd: acc per kick, phi for the energies), fp64 and fp32, on the reference's C backend.

At this size the reference's single-threaded Python driver needs ~45 s per force evaluation,
so the run is driven by oracle/integrators.py (pinned bit-for-bit to the reference's
integrators by tests/test_oracle_integrators.py) calling the UNMODIFIED reference C backend
(oracle/_ref) on contiguous i-slices from a thread pool -- the same arithmetic per particle as
the reference's own loop (each i is summed over all j sequentially).

    python tests/golden/make_golden_config2.py [threads]

Stores the energies before/after, the step count and the final state of 512 evenly spaced
particles (the inputs are regenerated in the test from tupan_b200.ics.make_plummer(seed=1))."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import integrators as oi  # noqa: E402
from tupan_b200 import ics  # noqa: E402

N, ETA, T_END, METHOD = 65536, 1.0 / 256, 1.0 / 64, "sia21s.dkd"
VEC = ("rx", "ry", "rz", "vx", "vy", "vz")


def main():
    threads = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    flat = {}
    for prec in ("float64", "float32"):
        src = ics.make_plummer(N, seed=1, dtype=prec)
        ins = {k: getattr(src, k).copy() for k in ("mass", "eps2") + VEC}
        t0 = time.time()
        b0 = oi.Bodies(ins, prec, "ref", threads)
        ke0, pe0 = b0.kinetic_energy, b0.potential_energy
        ps, steps = oi.evolve(ins, prec, METHOD, ETA, T_END, kind="ref", threads=threads)
        ke1, pe1 = ps.kinetic_energy, ps.potential_energy
        idx = np.linspace(0, N - 1, 512).astype(np.int64)
        for k in VEC:
            flat["%s/out/%s" % (prec, k)] = ps.a[k][idx]
        flat["%s/idx" % prec] = idx
        flat["%s/meta" % prec] = np.array([ETA, T_END, steps, float(ps.clock[0]), ke0, pe0, ke1, pe1])
        print(prec, "steps", steps, "eerr", ((ke1 + pe1) - (ke0 + pe0)) / (-pe1), "%.0fs" % (time.time() - t0),
              flush=True)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "config2_leapfrog_n65536.npz"),
                        **flat)


if __name__ == "__main__":
    main()
