"""Shared helpers for the parity tests: kernel table, argument marshalling, error metric."""
import numpy as np

import oracle

S8 = ("mass", "rx", "ry", "rz", "eps2", "vx", "vy", "vz")
S5 = ("mass", "rx", "ry", "rz", "eps2")
S14 = S8 + ("ax", "ay", "az", "jx", "jy", "jz")
SV = ("mass", "vx", "vy", "vz", "ax", "ay", "az")
CL = 128.0
PN7 = (7,) + tuple((1.0 / CL) ** k for k in range(1, 8))

# name -> (input attributes, default scalars, output vector groups)
KERNELS = {
    "phi_kernel": (S5, (), [(0,)]),
    "acc_kernel": (S5, (), [(0, 1, 2)]),
    "acc_jerk_kernel": (S8, (), [(0, 1, 2), (3, 4, 5)]),
    "snap_crackle_kernel": (S14, (), [(0, 1, 2), (3, 4, 5)]),
    "tstep_kernel": (S8, (1.0 / 64,), [(0,), (1,)]),
    "pnacc_kernel": (S8, PN7, [(0, 1, 2)]),
    "nreg_Xkernel": (S8, (1.0 / 64,), [(0, 1, 2), (3, 4, 5), (6,)]),
    "nreg_Vkernel": (SV, (1.0 / 64,), [(0, 1, 2), (3,)]),
    "sakura_kernel": (S8, (1.0 / 64, 1), [(0, 1, 2), (3, 4, 5)]),
}


def pn_scalars(order, c=CL):
    return (order,) + tuple((1.0 / c) ** k for k in range(1, 8))


def as_dict(ps, dtype):
    """ParticleSystem (or dict) -> dict of contiguous arrays of `dtype`; missing a/j arrays
    (needed by snap_crackle / nreg_V) are synthesised deterministically."""
    dtype = np.dtype(dtype)
    src = ps if isinstance(ps, dict) else ps.arrays()
    d = {k: np.ascontiguousarray(v, dtype) for k, v in src.items() if v.dtype.kind == "f"}
    n = len(d["mass"])
    rng = np.random.default_rng(12345)
    for k in ("ax", "ay", "az", "jx", "jy", "jz"):
        if k not in d:
            d[k] = rng.standard_normal(n).astype(dtype)
    return d


def run(lib, name, prec, I, J, scalars=None, ni=None, nj=None):
    """Call `name` from any library with the libtupan.h ABI; returns the output arrays."""
    attrs, default, _ = KERNELS[name]
    scalars = default if scalars is None else scalars
    ni = len(I["mass"]) if ni is None else ni
    nj = len(J["mass"]) if nj is None else nj
    dt = np.dtype(prec)
    outs = [np.full(ni, np.nan, dt) for _ in range(oracle.n_outputs(name))]
    args = ([ni] + [np.ascontiguousarray(I[a][:ni]) for a in attrs]
            + [nj] + [np.ascontiguousarray(J[a][:nj]) for a in attrs] + list(scalars) + outs)
    oracle.call(lib, name, prec, *args)
    return outs


def state_floors(I):
    """Denominator floors for sakura's (dr, dv) outputs: the rms size of the positions and
    velocities they are added to (integrator/sakura.py:31-36).  dr and dv are differences of
    nearly equal states (evolved minus initial two-body state, minus the free drift), so their
    own norm can be ~1e-9 of the operands and rounding noise of the REFERENCE is already
    >> 1e-12 relative to it; the meaningful scale is the state being updated."""
    r = np.sqrt(np.mean(I["rx"].astype(np.float64) ** 2 + I["ry"].astype(np.float64) ** 2
                        + I["rz"].astype(np.float64) ** 2))
    v = np.sqrt(np.mean(I["vx"].astype(np.float64) ** 2 + I["vy"].astype(np.float64) ** 2
                        + I["vz"].astype(np.float64) ** 2))
    return (float(r), float(v))


def rel_err(name, got, ref, floors=None):
    """max over particles and output groups of ||got - ref||_2 / max(||ref||_2, floor) (per
    particle); `floors` (one per output group) defaults to 0."""
    worst = 0.0
    for gi, grp in enumerate(KERNELS[name][2]):
        g = np.stack([got[k] for k in grp]).astype(np.float64)
        r = np.stack([ref[k] for k in grp]).astype(np.float64)
        if not np.all(np.isfinite(g) == np.isfinite(r)):
            return np.inf
        fin = np.all(np.isfinite(r), axis=0)
        d = np.sqrt(((g - r)[:, fin] ** 2).sum(0))
        nr = np.sqrt((r[:, fin] ** 2).sum(0))
        if floors is not None:
            nr = np.maximum(nr, floors[gi])
        scale = np.where(nr > 0, nr, 1.0)
        e = d / scale
        if e.size:
            worst = max(worst, float(e.max()))
    return worst


def cuda_lib(prec):
    from tupan_b200 import backend
    return backend.require_gpu(prec)


def cuda_run(name, prec, I, J, scalars=None, ni=None, nj=None):
    from tupan_b200 import backend
    lib = cuda_lib(prec)
    out = run(lib, name, prec, I, J, scalars, ni, nj)
    backend.check(lib, name)
    return out
