"""Generate tests/golden/integrators_fp{64,32}.npz by running the UNMODIFIED reference
integrators (ggf84/tupan, `tupan.integrator.Integrator`) on its own C backend.

Run in the build container only (needs /root/reference, gcc, cffi, scipy):

    python tests/golden/make_golden_integrators.py             # fp64
    python tests/golden/make_golden_integrators.py --use_sp    # fp32 (precision is read from
                                                               # sys.argv at import, ctype.py:13)
    python tests/golden/make_golden_integrators.py --config1   # adds BASELINE.json configs[0]:
                                                               # Plummer N=1024, (a)hermite4, t_end=1
    python tests/golden/make_golden_integrators.py --pn_order  # post-Newtonian SIA runs (fp64)

Each case stores the initial state, the state after the reference's own driver loop
(`while abs(time) < t_end: evolve_step(t_end)`, simulation.py:187-201), the number of steps,
the final time and the energies computed by the reference (`kinetic_energy`,
`potential_energy`, particles/body.py:262-306), so the tests need neither the reference nor
this script.
"""
import os
import sys
import tempfile
import time

os.environ["HOME"] = tempfile.mkdtemp(prefix="tupan_home_")   # ~/.tupan/cffi-cache-* must be writable
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
from tupan.lib.utils import ctype  # noqa: E402
from tupan.ics.plummer import make_plummer  # noqa: E402
from tupan.integrator import Integrator  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TAG = "fp32" if ctype.use_sp else "fp64"
STATE = ("mass", "eps2", "rx", "ry", "rz", "vx", "vy", "vz")
OUT = ("rx", "ry", "rz", "vx", "vy", "vz", "time", "tstep", "nstep")

SMALL = [
    # (method, n, eta, t_end)
    ("hermite2", 32, 1.0 / 64, 0.125), ("hermite4", 32, 1.0 / 64, 0.125),
    ("hermite6", 32, 1.0 / 64, 0.125), ("hermite8", 32, 1.0 / 64, 0.125),
    ("ahermite2", 32, 1.0 / 16, 0.0625), ("ahermite4", 32, 1.0 / 16, 0.0625),
    ("ahermite6", 32, 1.0 / 16, 0.0625), ("ahermite8", 32, 1.0 / 16, 0.0625),
    ("hermite4", 33, 1.0 / 50, 0.11),            # t_end not a multiple of eta: last step is cut
    ("sia21s.dkd", 32, 1.0 / 64, 0.125), ("sia21s.kdk", 32, 1.0 / 64, 0.125),
    ("sia21a.dkd", 32, 1.0 / 16, 0.0625), ("sia21a.kdk", 32, 1.0 / 16, 0.0625),
    ("sia22s.dkd", 32, 1.0 / 64, 0.125), ("sia22a.kdk", 32, 1.0 / 16, 0.0625),
    ("sia43s.kdk", 32, 1.0 / 64, 0.125), ("sia43a.dkd", 32, 1.0 / 16, 0.0625),
    ("sia44s.dkd", 32, 1.0 / 64, 0.125), ("sia45s.kdk", 32, 1.0 / 64, 0.125),
    ("sia46s.dkd", 32, 1.0 / 64, 0.125), ("sia67s.kdk", 32, 1.0 / 64, 0.125),
    ("sia69s.dkd", 32, 1.0 / 64, 0.125),
    ("sia21h.dkd", 32, 1.0 / 16, 0.0625), ("sia21h.kdk", 32, 1.0 / 16, 0.0625),
    ("sia43h.kdk", 32, 1.0 / 16, 0.0625),
    ("nreg", 32, 1.0 / 64, 0.125), ("anreg", 32, 1.0 / 64, 0.125),
    ("sakura", 32, 1.0 / 64, 0.125), ("asakura", 32, 1.0 / 16, 0.0625),
]
CONFIG1 = [("hermite4", 1024, 1.0 / 64, 1.0), ("ahermite4", 1024, 1.0 / 64, 1.0)]
# post-Newtonian SIA (kick_pn / drift_pn, sia.py:90-159): needs "--pn_order" in sys.argv at import
# (particles/body.py:532); (method, n, eta, t_end, pn_order, clight).  kdk only: the dkd variants
# touch pn_mvx before it is registered unless an earlier run in the same process created it.
PN = [("sia21s.kdk", 32, 1.0 / 64, 0.0625, 7, 8.0), ("sia21a.kdk", 32, 1.0 / 16, 1.0 / 32, 7, 8.0),
      ("sia43s.kdk", 32, 1.0 / 64, 0.0625, 7, 8.0), ("sia22a.kdk", 32, 1.0 / 16, 1.0 / 32, 4, 16.0),
      ("sia21s.kdk", 32, 1.0 / 64, 0.0625, 2, 4.0)]
PN_OUT = ("wx", "wy", "wz", "pn_ke", "pn_mrx", "pn_mry", "pn_mrz", "pn_mvx", "pn_mvy", "pn_mvz",
          "pn_amx", "pn_amy", "pn_amz")


def run_case(method, n, eta, t_end, seed=1, pn_order=0, clight=None):
    ps = make_plummer(n, 4.0 / n, ("equalmass",), seed=seed)
    rec = {"in/" + k: np.array(getattr(ps, k)) for k in STATE}
    type(ps).include_pn_corrections = False      # energies of the initial state are Newtonian
    ke0, pe0 = ps.kinetic_energy, ps.potential_energy
    if pn_order:
        it = Integrator(eta, 0.0, ps, method=method, pn_order=pn_order, clight=clight)
    else:
        it = Integrator(eta, 0.0, ps, method=method)
    steps = 0
    t0 = time.time()
    while abs(it.time) < t_end:
        it.evolve_step(t_end)
        steps += 1
    ps = it.particle_system
    for k in OUT + ("id",) + (PN_OUT if pn_order else ()):
        rec["out/" + k] = np.array(getattr(ps, k))
    ke1, pe1 = ps.kinetic_energy, ps.potential_energy            # ke includes pn_ke when PN is on
    rec["meta"] = np.array([eta, t_end, steps, it.time, ke0, pe0, ke1, pe1, pn_order, clight or 0.0],
                           dtype=np.float64)
    print("%-4s %-12s n=%-5d steps=%-6d t=%.6f eerr=%+.3e  %.1fs" % (
        TAG, method, n, steps, it.time, ((ke1 + pe1) - (ke0 + pe0)) / (-pe1), time.time() - t0), flush=True)
    return rec


# The reference's own `nreg` does not terminate in fp32 (its driver loop approaches t_end in
# ever smaller fictitious-time steps and never reaches it); probed here with a 30 s timeout.
SKIP_FP32 = ("nreg",)


def main():
    cases = [c for c in SMALL if not (ctype.use_sp and c[0] in SKIP_FP32)]
    name = "integrators"
    if "--config1" in sys.argv:
        cases = list(CONFIG1)
        name = "integrators_config1"
    if "--pn_order" in sys.argv:
        cases = list(PN)
        name = "integrators_pn"
    flat = {}
    for case in cases:
        method, n, eta, t_end = case[:4]
        rec = run_case(method, n, eta, t_end, 1, *case[4:])
        key = "%s_n%d" % (method, n) + ("_pn%d_c%g" % case[4:] if len(case) > 4 else "")
        for k, v in rec.items():
            flat[key + "/" + k] = v
    path = os.path.join(HERE, "%s_%s.npz" % (name, TAG))
    np.savez_compressed(path, **flat)
    print(path, len(cases), "cases")


if __name__ == "__main__":
    main()
