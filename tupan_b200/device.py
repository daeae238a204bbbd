"""Device-resident entry points (Part 2 of include/libtupan_cuda.h) for torch tensors.

PyTorch is plumbing here: it owns device memory and streams; every kernel launched is one
of ours.  Particle state is a dict of 1-D CUDA tensors (SoA, same attribute names as the
host containers); outputs are written into caller-provided tensors, asynchronously on the
current torch stream.
"""
import ctypes

import torch

from . import backend

KERNEL_INPUTS = {
    "phi_kernel": ("mass", "rx", "ry", "rz", "eps2"),
    "acc_kernel": ("mass", "rx", "ry", "rz", "eps2"),
    "acc_jerk_kernel": ("mass", "rx", "ry", "rz", "eps2", "vx", "vy", "vz"),
    "snap_crackle_kernel": ("mass", "rx", "ry", "rz", "eps2", "vx", "vy", "vz",
                            "ax", "ay", "az", "jx", "jy", "jz"),
    "tstep_kernel": ("mass", "rx", "ry", "rz", "eps2", "vx", "vy", "vz"),
    "pnacc_kernel": ("mass", "rx", "ry", "rz", "eps2", "vx", "vy", "vz"),
    "nreg_Xkernel": ("mass", "rx", "ry", "rz", "eps2", "vx", "vy", "vz"),
    "nreg_Vkernel": ("mass", "vx", "vy", "vz", "ax", "ay", "az"),
    "sakura_kernel": ("mass", "rx", "ry", "rz", "eps2", "vx", "vy", "vz"),
}
KERNEL_OUTPUTS = {
    "phi_kernel": ("phi",),
    "acc_kernel": ("ax", "ay", "az"),
    "acc_jerk_kernel": ("ax", "ay", "az", "jx", "jy", "jz"),
    "snap_crackle_kernel": ("sx", "sy", "sz", "cx", "cy", "cz"),
    "tstep_kernel": ("tstep", "tstepij"),
    "pnacc_kernel": ("pnax", "pnay", "pnaz"),
    "nreg_Xkernel": ("mrx", "mry", "mrz", "ax", "ay", "az", "u"),
    "nreg_Vkernel": ("mvx", "mvy", "mvz", "mk"),
    "sakura_kernel": ("drx", "dry", "drz", "dvx", "dvy", "dvz"),
}


def prec_of_tensor(t):
    if t.dtype == torch.float64:
        return "float64"
    if t.dtype == torch.float32:
        return "float32"
    raise TypeError("float32 / float64 tensors only")


def ptr_array(tensors):
    for t in tensors:
        if not (t.is_cuda and t.is_contiguous()):
            raise TypeError("contiguous CUDA tensors required")
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def scal_array(scalars):
    s = list(scalars) + [0.0] * (8 - len(scalars))
    return (ctypes.c_double * len(s))(*[float(x) for x in s])


def current_stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_device(ps, device="cuda", dtype=None):
    """Host particle container (or dict of numpy arrays) -> dict of CUDA tensors."""
    src = ps if isinstance(ps, dict) else ps.arrays()
    out = {}
    for k, v in src.items():
        if v.dtype.kind != "f":
            continue
        t = torch.from_numpy(v)
        out[k] = t.to(device=device, dtype=dtype or t.dtype, non_blocking=False).contiguous()
    return out


def run(kernel, ips, jps, scalars=(), out=None):
    """out[attr] = kernel(ips, jps) on the current torch stream; returns the output dict."""
    ins = KERNEL_INPUTS[kernel]
    it = [ips[a] for a in ins]
    jt = [jps[a] for a in ins]
    prec = prec_of_tensor(it[0])
    lib = backend.require_gpu(prec)
    ni, nj = it[0].numel(), jt[0].numel()
    if out is None:
        out = {a: torch.empty(ni, dtype=it[0].dtype, device=it[0].device) for a in KERNEL_OUTPUTS[kernel]}
    ot = [out[a] for a in KERNEL_OUTPUTS[kernel]]
    rc = lib.tupan_cuda_run_dev(backend.KERNEL_IDS[kernel], ni, ptr_array(it), nj, ptr_array(jt),
                                scal_array(scalars), ptr_array(ot), current_stream())
    if rc != 0:
        backend.check(lib, kernel)
        raise backend.TupanCudaError("%s failed with code %d" % (kernel, rc))
    return out


def fma_peak(prec="float64", ms=200.0):
    """Sustained FMA-pipe TFLOP/s of this board in `prec`, and the SM clock that implies."""
    lib = backend.require_gpu(prec)
    tf, mhz = ctypes.c_double(), ctypes.c_double()
    rc = lib.tupan_cuda_fma_peak(ms, ctypes.byref(tf), ctypes.byref(mhz))
    if rc != 0:
        backend.check(lib, "fma_peak")
    return tf.value, mhz.value
