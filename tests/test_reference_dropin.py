"""Drop-in seam against the real, UNMODIFIED reference (ggf84/tupan).

The reference is found as an installed copy under ``baseline/_ref`` (``pip install --target``
of /root/reference, done by ``__graft_entry__.build()`` in the build container; git-ignored, it
travels to the GPU box with the snapshot) or, in the build container, as /root/reference itself.

``tupan_b200.extensions.install()`` swaps the CUDA kernel objects into the reference's
``tupan.lib.extensions``; the reference's own ``ParticleSystem.set_*`` force setters and its own
``tupan.integrator.Integrator`` must then reach our ``CUDAKernel`` with the reference's argument
marshalling (extensions.py:63-97,654-666; particles/body.py:324-361).

* without a GPU the call has to end in TupanCudaError (no CPU fallback) -- CPU test;
* on the B200 the reference's integrators run to ``t_end`` on the CUDA kernels and must reproduce
  their own C-backend runs: BASELINE.json configs[0] (Plummer N = 1024, ahermite4, eta = 1/64,
  t_end = 1) against the golden run of the C backend (tests/golden/integrators_config1_fp64.npz,
  18 379 adaptive steps), and shorter runs of sia21s.dkd, sia21a.kdk, ahermite6, hermite8,
  asakura, sakura, sia21h.dkd (Kepler leaves) and anreg against the C backend run in the same
  process just before the swap.  Stated tolerances: same number of steps and final time; relative
  energy error within 1e-10 of the C backend's; states within 1e-9 (1e-7 after 18 379 steps).
"""
import json
import os
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.path.join(ROOT, "baseline", "_ref"), "/root/reference")
REF = next((p for p in CANDIDATES if os.path.isdir(os.path.join(p, "tupan", "lib", "src"))), None)

pytestmark = pytest.mark.skipif(REF is None, reason="no reference (baseline/_ref or /root/reference)")

SEAM = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(ref)r)
import numpy as np
import torch
import tupan.lib.extensions as ref_ext
from tupan.ics.plummer import make_plummer
from tupan.integrator.hermite import Hermite
import tupan_b200.extensions as ours
from tupan_b200.backend import CUDAKernel, TupanCudaError

ps = make_plummer(32, 4.0 / 32, ("equalmass",), seed=1)
ref = [np.array(a) for a in ref_ext.acc_jerk.calc(ps, ps)]        # reference C backend first

ours.install(ref_ext)
assert ref_ext.backend == "CUDA"
for name in ("phi", "acc", "acc_jerk", "snap_crackle", "tstep", "pnacc", "sakura", "nreg_x", "nreg_v", "kepler"):
    assert isinstance(getattr(ref_ext, name).kernel, CUDAKernel), name
assert ours.clight is ref_ext.clight

# the reference's own setter now marshals into our adapter (same array objects, same order)
ext = ref_ext.acc_jerk
ext.set_args(ps, ps)
assert ext._inargs[0] == ps.n and ext._inargs[9] == ps.n
assert ext._inargs[1] is ps.mass and ext._inargs[8] is ps.vz
assert len(ext.kernel.args) == 24

if torch.cuda.is_available():
    ps.set_acc_jerk(ps)
    got = [ps.ax, ps.ay, ps.az, ps.jx, ps.jy, ps.jz]
    for g, r in zip(got, ref):
        assert np.allclose(g, r, rtol=1e-11, atol=0)
    print("DROPIN-GPU-OK")
else:
    try:
        ps.set_acc_jerk(ps)
    except TupanCudaError as e:
        print("DROPIN-NOGPU-OK", str(e)[:60])
    else:
        raise SystemExit("a CPU fallback answered")
    # an unmodified integrator reaches the same seam
    try:
        Hermite(1.0 / 64, 0.0, ps, method="hermite4").evolve_step(1.0 / 8)
    except TupanCudaError:
        print("INTEGRATOR-SEAM-OK")
'''

# The reference's integrators, unmodified, first on the reference's C backend, then -- after
# install() -- on the CUDA kernels; one JSON line per case.
DRIVE = r'''
import json, os, sys, time
sys.path.insert(0, %(root)r); sys.path.insert(0, %(ref)r)
import numpy as np
import tupan.lib.extensions as ref_ext
from tupan.ics.plummer import make_plummer
from tupan.integrator import Integrator
import tupan_b200.extensions as ours
from tupan_b200 import backend

VEC = ("rx", "ry", "rz", "vx", "vy", "vz")

def run(method, n, eta, t_end):
    ps = make_plummer(n, 4.0 / n, ("equalmass",), seed=1)
    ke0, pe0 = ps.kinetic_energy, ps.potential_energy
    it = Integrator(eta, 0.0, ps, method=method)
    steps, t0 = 0, time.time()
    while abs(it.time) < t_end:
        it.evolve_step(t_end)
        steps += 1
    ps = it.particle_system
    ke1, pe1 = ps.kinetic_energy, ps.potential_energy
    return dict(steps=steps, time=float(it.time), eerr=float(((ke1 + pe1) - (ke0 + pe0)) / (-pe1)),
                ke0=float(ke0), pe0=float(pe0), state={k: np.array(getattr(ps, k)) for k in VEC},
                seconds=time.time() - t0)

def relmax(a, b):
    return float(max(np.max(np.abs(a[k] - b[k])) / np.max(np.abs(b[k])) for k in VEC))

CASES = [("sia21s.dkd", 1024, 1.0 / 64, 0.25), ("sia21a.kdk", 256, 1.0 / 16, 0.125),
         ("ahermite6", 256, 1.0 / 32, 0.125), ("hermite8", 128, 1.0 / 64, 0.0625),
         ("asakura", 128, 1.0 / 16, 0.0625), ("sakura", 64, 1.0 / 64, 0.0625),
         ("sia21h.dkd", 64, 1.0 / 16, 0.25), ("anreg", 32, 1.0 / 64, 0.0625)]
cpu = {c: run(*c) for c in CASES}

# call rate of the two backends through the SAME reference wrapper (extensions.AccJerk.calc)
def rate(n, reps):
    ps = make_plummer(n, 4.0 / n, ("equalmass",), seed=1)
    ref_ext.acc_jerk.calc(ps, ps)
    t0 = time.perf_counter()
    for _ in range(reps):
        ref_ext.acc_jerk.calc(ps, ps)
    return reps / (time.perf_counter() - t0)
c_rate = {n: rate(n, r) for n, r in ((256, 200), (1024, 40), (4096, 4))}

ours.install(ref_ext)
assert ref_ext.backend == "CUDA"
lib = backend.require_gpu("float64")
launches0 = lib.tupan_cuda_launch_count()
ok = True
for c in CASES:
    g = run(*c)
    line = dict(case="%%s n=%%d eta=1/%%g t_end=%%g" %% (c[0], c[1], 1 / c[2], c[3]), steps=g["steps"], steps_c=cpu[c]["steps"],
                time=g["time"], time_c=cpu[c]["time"], eerr=g["eerr"], eerr_c=cpu[c]["eerr"],
                state_relmax=relmax(g["state"], cpu[c]["state"]), seconds_cuda=g["seconds"],
                seconds_c=cpu[c]["seconds"])
    line["ok"] = bool(line["steps"] == line["steps_c"] and line["time"] == line["time_c"]
                      and abs(line["eerr"] - line["eerr_c"]) <= 1e-10 and line["state_relmax"] <= 1e-9)
    ok = ok and line["ok"]
    print("CASE " + json.dumps(line), flush=True)

# BASELINE.json configs[0] against the golden C-backend run
z = np.load(os.path.join(%(root)r, "tests", "golden", "integrators_config1_fp64.npz"))
eta, t_end, steps_ref, t_ref, ke0r, pe0r, ke1r, pe1r = z["ahermite4_n1024/meta"][:8]
g = run("ahermite4", 1024, float(eta), float(t_end))
eerr_ref = ((ke1r + pe1r) - (ke0r + pe0r)) / (-pe1r)
gold = {k: z["ahermite4_n1024/out/" + k] for k in VEC}
line = dict(case="config0: ahermite4 n=1024 eta=1/64 t_end=1 (golden C-backend run)", steps=g["steps"],
            steps_c=int(steps_ref), time=g["time"], time_c=float(t_ref), eerr=g["eerr"], eerr_c=float(eerr_ref),
            state_relmax=relmax(g["state"], gold), seconds_cuda=g["seconds"],
            ke0_match=abs(g["ke0"] / ke0r - 1), pe0_match=abs(g["pe0"] / pe0r - 1))
line["ok"] = bool(line["steps"] == line["steps_c"] and line["time"] == line["time_c"]
                  and abs(line["eerr"] - line["eerr_c"]) <= 1e-10 and line["state_relmax"] <= 1e-7)
ok = ok and line["ok"]
print("CASE " + json.dumps(line), flush=True)

g_rate = {n: rate(n, r) for n, r in ((256, 2000), (1024, 2000), (4096, 500))}
print("RATE " + json.dumps(dict(what="extensions.AccJerk.calc(ps, ps) calls per second, reference wrapper, host arrays",
                                c_backend=c_rate, cuda_backend=g_rate)), flush=True)
print("LAUNCHES", lib.tupan_cuda_launch_count() - launches0)
print("DRIVE-GPU-OK" if ok else "DRIVE-GPU-MISMATCH")
'''


def _run(script, timeout):
    home = tempfile.mkdtemp(prefix="tupan_home_")        # ~/.tupan/cffi-cache-* must be writable
    env = dict(os.environ, HOME=home)
    return subprocess.run([sys.executable, "-c", script % {"root": ROOT, "ref": REF}], env=env,
                          capture_output=True, text=True, timeout=timeout)


def test_install_into_unmodified_reference():
    p = _run(SEAM, 600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "DROPIN-GPU-OK" in p.stdout or ("DROPIN-NOGPU-OK" in p.stdout and "INTEGRATOR-SEAM-OK" in p.stdout), \
        p.stdout + p.stderr


@pytest.mark.gpu
def test_reference_integrators_drive_the_cuda_kernels():
    p = _run(DRIVE, 3000)
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):                                # keep the evidence (copied to profiles/ per round)
        with open(os.path.join(out, "dropin_gpu.jsonl"), "w") as f:
            f.write("\n".join(l for l in p.stdout.splitlines() if l.startswith(("CASE", "RATE", "LAUNCHES"))) + "\n")
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "DRIVE-GPU-OK" in p.stdout, p.stdout[-4000:]
    cases = [json.loads(l[5:]) for l in p.stdout.splitlines() if l.startswith("CASE ")]
    assert len(cases) == 9 and all(c["ok"] for c in cases), cases
    launches = [int(l.split()[1]) for l in p.stdout.splitlines() if l.startswith("LAUNCHES")]
    assert launches and launches[0] > 10000               # the integrators really ran on our kernels
