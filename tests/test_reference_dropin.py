"""Drop-in seam against the real reference checkout (only where /root/reference exists, i.e.
the build container; the GPU box has no reference and skips this file).

`tupan_b200.extensions.install()` swaps the CUDA kernel objects into the UNMODIFIED
`tupan.lib.extensions`; the reference's own `ParticleSystem.set_*` force setters and its
integrators must then reach our `CUDAKernel` with the reference's argument marshalling.  In a
container without a GPU the call has to end in TupanCudaError (no CPU fallback)."""
import os
import subprocess
import sys
import tempfile

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tupan")), reason="no reference checkout")

SCRIPT = r'''
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


def test_install_into_unmodified_reference():
    home = tempfile.mkdtemp(prefix="tupan_home_")        # ~/.tupan/cffi-cache-* must be writable
    env = dict(os.environ, HOME=home)
    p = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "ref": REF}], env=env,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "DROPIN-GPU-OK" in p.stdout or ("DROPIN-NOGPU-OK" in p.stdout and "INTEGRATOR-SEAM-OK" in p.stdout), \
        p.stdout + p.stderr
