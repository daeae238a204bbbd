"""ctypes binding of ``libtupan_cuda_fp{32,64}.so`` and the ``CUDAKernel`` adapter.

``CUDAKernel(prec, name)`` implements the kernel-adapter protocol that the reference's
``tupan/lib/extensions.py:63-97`` expects from ``get_kernel`` (reference implementations:
``cffi_backend.py:90-126`` ``CKernel`` and ``opencl_backend.py:104-171`` ``CLKernel``):
``.cty`` converters, ``set_gsize``, ``set_args``, ``run``, ``map_buffers``.  The adapter calls
the Part-1 entry points of ``include/libtupan_cuda.h`` (host pointers in, host pointers out,
synchronous), so numpy-owned particle arrays work unchanged.

There is no CPU fallback: if the library is missing or a CUDA call fails, this module raises.
"""
import ctypes
import os
from collections import namedtuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBDIR = os.path.join(HERE, "lib")

KERNEL_IDS = {
    "phi_kernel": 0, "acc_kernel": 1, "acc_jerk_kernel": 2, "snap_crackle_kernel": 3,
    "tstep_kernel": 4, "pnacc_kernel": 5, "nreg_Xkernel": 6, "nreg_Vkernel": 7,
    "sakura_kernel": 8, "kepler_solver_kernel": 9,
}

# Argument kinds in libtupan.h order: n = UINT count, P = const REAL* in, O = REAL* out,
# r = REAL scalar, u = UINT scalar, i = INT scalar.
_I5, _I8, _I14, _I7 = "nPPPPP", "nPPPPPPPP", "nPPPPPPPPPPPPPP", "nPPPPPPP"
SIGNATURES = {
    "phi_kernel": _I5 + _I5 + "O",
    "acc_kernel": _I5 + _I5 + "OOO",
    "acc_jerk_kernel": _I8 + _I8 + "OOOOOO",
    "snap_crackle_kernel": _I14 + _I14 + "OOOOOO",
    "tstep_kernel": _I8 + _I8 + "r" + "OO",
    "pnacc_kernel": _I8 + _I8 + "urrrrrrr" + "OOO",
    "nreg_Xkernel": _I8 + _I8 + "r" + "OOOOOOO",
    "nreg_Vkernel": _I7 + _I7 + "r" + "OOOO",
    "sakura_kernel": _I8 + _I8 + "ri" + "OOOOOO",
    "kepler_solver_kernel": "PPPPPPPP" + "r" + "OOOOOO",
}

PART2 = {
    # name: (restype, argtypes)
    "tupan_cuda_info": (ctypes.c_int, [ctypes.c_int] + [ctypes.POINTER(ctypes.c_int)] * 4),
    "tupan_cuda_run_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "tupan_cuda_kepler_dev": (ctypes.c_int, [ctypes.c_longlong, ctypes.c_void_p, ctypes.c_double,
                                             ctypes.c_void_p, ctypes.c_void_p]),
    "tupan_cuda_kepler_limit_hits": (ctypes.c_longlong, []),
    "tupan_cuda_kepler_cleanup_pairs": (ctypes.c_longlong, []),
    "tupan_cuda_row_width": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p]),
    "tupan_cuda_n_acc": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p]),
    "tupan_cuda_pack_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p]),
    "tupan_cuda_sweep_slots": (ctypes.c_int, [ctypes.c_int, ctypes.c_longlong, ctypes.c_longlong,
                                              ctypes.c_void_p]),
    "tupan_cuda_sweep_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "tupan_cuda_finalize_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p]),
    "tupan_cuda_abs_min_dev": (ctypes.c_int, [ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p]),
    "tupan_cuda_peer_alloc": (ctypes.c_int, [ctypes.c_longlong, ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]),
    "tupan_cuda_peer_open": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "tupan_cuda_peer_close": (ctypes.c_int, [ctypes.c_void_p]),
    "tupan_cuda_peer_free": (ctypes.c_int, [ctypes.c_void_p]),
    "tupan_cuda_peer_barrier_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                                   ctypes.c_void_p]),
    "tupan_cuda_peer_timeouts": (ctypes.c_longlong, []),
    "tupan_cuda_sweep_multi_slots": (ctypes.c_int, [ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p,
                                                    ctypes.c_void_p]),
    "tupan_cuda_sweep_multi_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_int,
                                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_int, ctypes.c_void_p]),
    "tupan_cuda_init": (ctypes.c_int, []),
    "tupan_cuda_last_error": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
    "tupan_cuda_clear_error": (None, []),
    "tupan_cuda_force_plan": (None, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "tupan_cuda_last_plan": (None, [ctypes.POINTER(ctypes.c_int)] * 3),
    "tupan_cuda_plan_query": (ctypes.c_int, [ctypes.c_int, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p]
                              + [ctypes.POINTER(ctypes.c_int)] * 3),
    "tupan_cuda_set_timing": (None, [ctypes.c_int]),
    "tupan_cuda_last_times": (None, [ctypes.POINTER(ctypes.c_float)] * 5),
    "tupan_cuda_sum_times": (ctypes.c_int, [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_longlong)]),
    "tupan_cuda_launch_count": (ctypes.c_longlong, []),
    "tupan_cuda_count_launches": (None, [ctypes.c_longlong]),
    "tupan_cuda_sm_count": (ctypes.c_int, []),
    "tupan_cuda_fma_peak": (ctypes.c_int, [ctypes.c_double, ctypes.POINTER(ctypes.c_double),
                                           ctypes.POINTER(ctypes.c_double)]),
    "tupan_cuda_pipe_probe": (ctypes.c_int, [ctypes.c_int, ctypes.c_double, ctypes.POINTER(ctypes.c_double)]),
    "tupan_cuda_real_bytes": (ctypes.c_int, []),
    # Part 3: O(N) integrator updates on device-resident state
    "tupan_cuda_step_begin_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "tupan_cuda_block_predict_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
                                                    ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]),
    "tupan_cuda_block_correct_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
                                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                    ctypes.c_void_p]),
    "tupan_cuda_block_select_dev": (ctypes.c_int, [ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_void_p, ctypes.c_void_p]),
    "tupan_cuda_block_quantize_dev": (ctypes.c_int, [ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
                                                     ctypes.c_double, ctypes.c_double, ctypes.c_void_p,
                                                     ctypes.c_void_p, ctypes.c_void_p]),
    "tupan_cuda_hermite_predict_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p,
                                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                      ctypes.c_void_p]),
    "tupan_cuda_hermite_correct_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p,
                                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                      ctypes.c_void_p, ctypes.c_void_p]),
    "tupan_cuda_axpy_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_double, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]),
    "tupan_cuda_scale_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_double, ctypes.c_void_p]),
    "tupan_cuda_step_end_dev": (ctypes.c_int, [ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "tupan_cuda_pn_kick_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_double,
                                              ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]),
    "tupan_cuda_stamp_dev": (ctypes.c_int, [ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_double, ctypes.c_void_p]),
    "tupan_cuda_reduce_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_double,
                                             ctypes.c_void_p, ctypes.c_void_p]),
}

_PREC = {
    "float64": dict(real=ctypes.c_double, uint=ctypes.c_ulong, int=ctypes.c_long, np=np.dtype(np.float64),
                    npuint=np.dtype(np.uint64), npint=np.dtype(np.int64), tag="fp64"),
    "float32": dict(real=ctypes.c_float, uint=ctypes.c_uint, int=ctypes.c_int, np=np.dtype(np.float32),
                    npuint=np.dtype(np.uint32), npint=np.dtype(np.int32), tag="fp32"),
}


class TupanCudaError(RuntimeError):
    pass


def prec_of(dtype_or_name):
    if isinstance(dtype_or_name, str) and dtype_or_name in _PREC:
        return dtype_or_name
    dt = np.dtype(dtype_or_name)
    if dt == np.float64:
        return "float64"
    if dt == np.float32:
        return "float32"
    raise TypeError("tupan kernels exist for float32 and float64 only, not %s" % dt)


def lib_path(prec):
    return os.path.join(LIBDIR, "libtupan_cuda_%s.so" % _PREC[prec]["tag"])


_libs = {}


def load(prec="float64"):
    """Load (once) and bind the library of the given precision.  Raises if it is not built."""
    prec = prec_of(prec)
    if prec in _libs:
        return _libs[prec]
    path = lib_path(prec)
    if not os.path.exists(path):
        raise TupanCudaError(
            "%s is not built; run `python -m tupan_b200.build` (there is no CPU fallback)" % path)
    lib = ctypes.CDLL(path)
    t = _PREC[prec]
    kinds = {"n": t["uint"], "u": t["uint"], "i": t["int"], "r": t["real"],
             "P": ctypes.c_void_p, "O": ctypes.c_void_p}
    for name, sig in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = None
        fn.argtypes = [kinds[k] for k in sig]
    for name, (res, args) in PART2.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _libs[prec] = lib
    return lib


def check(lib, where=""):
    """Raise if the library recorded a failure (the Part-1 functions are void, like the
    reference's, so the error channel is tupan_cuda_last_error)."""
    buf = ctypes.create_string_buffer(256)
    code = lib.tupan_cuda_last_error(buf, 256)
    if code != 0:
        lib.tupan_cuda_clear_error()
        raise TupanCudaError("%s failed (%d): %s" % (where or "tupan_cuda", code, buf.value.decode()))


def require_gpu(prec="float64"):
    lib = load(prec)
    if lib.tupan_cuda_init() != 0:
        check(lib, "tupan_cuda_init")
    return lib


Types = namedtuple("Types", ["c_int", "c_int_p", "c_uint", "c_uint_p", "c_real", "c_real_p"])


class CUDAKernel(object):
    """Adapter with the protocol of the reference's CKernel / CLKernel."""

    def __init__(self, prec, name):
        self.prec = prec_of(prec)
        self.name = name
        self.lib = load(self.prec)
        self.kernel = getattr(self.lib, name)
        t = _PREC[self.prec]

        def real_p(x):
            if not (isinstance(x, np.ndarray) and x.dtype == t["np"] and x.flags.c_contiguous):
                raise TypeError("%s: expected a C-contiguous %s array" % (name, t["np"]))
            return x.ctypes.data

        def int_p(dt):
            def conv(x):
                if not (isinstance(x, np.ndarray) and x.dtype == dt and x.flags.c_contiguous):
                    raise TypeError("%s: expected a C-contiguous %s array" % (name, dt))
                return x.ctypes.data
            return conv

        self.cty = Types(c_int=int, c_int_p=int_p(t["npint"]),
                         c_uint=int, c_uint_p=int_p(t["npuint"]),
                         c_real=float, c_real_p=real_p)
        self.args = None
        self.global_size = None

    def set_gsize(self, ni, nj):
        # launch shape is chosen inside the library from (ni, nj); kept for protocol parity
        self.global_size = (int(ni), int(nj))

    def allocate_local_memory(self, numbufs, sctype):
        return []

    def set_args(self, args, start=0):
        self.args = list(args)

    def map_buffers(self, arrays, buffers):
        # outputs were written into the caller's arrays by run(); nothing to copy
        return arrays

    def run(self):
        self.kernel(*self.args)
        check(self.lib, self.name)
