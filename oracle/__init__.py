"""CPU oracle for the tupan pairwise kernels -- TEST INFRASTRUCTURE ONLY.

Two CPU implementations with the ABI of the reference's ``tupan/lib/src/libtupan.h:2-246``:

* ``kind="oracle"``: our plain-C restatement, ``oracle/tupan_oracle.c``.
* ``kind="ref"``:    the unmodified reference C backend compiled in place from
  ``/root/reference/tupan/lib/src`` by ``oracle/Makefile`` into ``oracle/_ref/``
  (git-ignored; travels to the GPU box as a prebuilt ``.so``).

Parity status: pinned -- ``tests/test_oracle.py`` checks the restatement bit-for-bit
against ``oracle/_ref`` and against ``tests/golden/*.npz`` produced by the reference's own
Python stack.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this package.  Nothing in ``tupan_b200/`` does, and the product path has no CPU fallback.
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# Argument kinds, in the order of libtupan.h.  'n' = UINT count, 'P' = const REAL* input,
# 'r' = REAL scalar, 'u' = UINT scalar, 'i' = INT scalar, 'O' = REAL* output.
# A ';' separates the i-block, the j-block, the scalars and the outputs for readability.
_I5, _I8, _I14, _I7 = "nPPPPP", "nPPPPPPPP", "nPPPPPPPPPPPPPP", "nPPPPPPP"
SIGNATURES = {
    "phi_kernel": _I5 + _I5 + "O",                       # libtupan.h:2-15
    "acc_kernel": _I5 + _I5 + "OOO",                     # :17-32
    "acc_jerk_kernel": _I8 + _I8 + "OOOOOO",             # :34-58
    "snap_crackle_kernel": _I14 + _I14 + "OOOOOO",       # :60-96
    "tstep_kernel": _I8 + _I8 + "r" + "OO",              # :98-119
    "pnacc_kernel": _I8 + _I8 + "urrrrrrr" + "OOO",      # :121-150
    "nreg_Xkernel": _I8 + _I8 + "r" + "OOOOOOO",         # :152-178
    "nreg_Vkernel": _I7 + _I7 + "r" + "OOOO",            # :180-201
    "sakura_kernel": _I8 + _I8 + "ri" + "OOOOOO",        # :203-229
    "kepler_solver_kernel": "PPPPPPPP" + "r" + "OOOOOO", # :231-246
}

_PREC = {
    "float64": dict(real=ctypes.c_double, uint=ctypes.c_ulong, int=ctypes.c_long,
                    np=np.float64, tag="fp64"),
    "float32": dict(real=ctypes.c_float, uint=ctypes.c_uint, int=ctypes.c_int,
                    np=np.float32, tag="fp32"),
    # x87 extended precision (numpy longdouble): the "truth" build of the restatement, oracle only
    "float128": dict(real=ctypes.c_longdouble, uint=ctypes.c_ulong, int=ctypes.c_long,
                     np=np.longdouble, tag="fp80"),
}


def lib_path(kind, prec):
    tag = _PREC[prec]["tag"]
    if kind == "oracle":
        return os.path.join(HERE, "libtupan_oracle_%s.so" % tag)
    if kind == "ref":
        return os.path.join(HERE, "_ref", "libtupan_ref_%s.so" % tag)
    raise ValueError(kind)


def build(quiet=True):
    """Compile the restatement and, when /root/reference is present, oracle/_ref."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def have(kind, prec="float64"):
    return os.path.exists(lib_path(kind, prec))


def bind(cdll, prec):
    """Attach argtypes for the ten libtupan.h entry points to an already loaded library."""
    t = _PREC[prec]
    rp = ctypes.c_void_p
    kinds = {"n": t["uint"], "u": t["uint"], "i": t["int"], "r": t["real"], "P": rp, "O": rp}
    for name, sig in SIGNATURES.items():
        fn = getattr(cdll, name)
        fn.restype = None
        fn.argtypes = [kinds[k] for k in sig]
    return cdll


_cache = {}


def load(kind="oracle", prec="float64"):
    key = (kind, prec)
    if key not in _cache:
        path = lib_path(kind, prec)
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        _cache[key] = bind(ctypes.CDLL(path), prec)
    return _cache[key]


def n_outputs(name):
    return SIGNATURES[name].count("O")


def _marshal(name, prec, args):
    """numpy arrays / python scalars -> ctypes values, checking dtype and contiguity."""
    sig = SIGNATURES[name]
    if len(args) != len(sig):
        raise TypeError("%s expects %d arguments, got %d" % (name, len(sig), len(args)))
    npdt = _PREC[prec]["np"]
    out = []
    for k, a in zip(sig, args):
        if k in "PO":
            if not (isinstance(a, np.ndarray) and a.dtype == npdt and a.flags.c_contiguous):
                raise TypeError("%s: array arguments must be C-contiguous %s" % (name, npdt))
            out.append(a.ctypes.data)
        elif k in "nui":
            out.append(int(a))
        else:
            out.append(float(a))
    return out


def call(lib, name, prec, *args):
    """Call ``name`` in ``lib`` (any library with the libtupan.h ABI) on numpy arrays."""
    getattr(lib, name)(*_marshal(name, prec, args))


def call_threaded(lib, name, prec, nthreads, *args):
    """Same result as :func:`call`, computed as ``nthreads`` contiguous i-slices in a thread
    pool.  Legal because every kernel is ``out[i] = reduce_j f(i, j)`` with ni != nj part of
    the API; ctypes releases the GIL during the call.  (Not for kepler_solver_kernel.)"""
    sig = SIGNATURES[name]
    ni = int(args[0])
    nthreads = max(1, min(nthreads, ni))
    first_j = sig.index("n", 1)
    bounds = np.linspace(0, ni, nthreads + 1).astype(np.int64)

    def piece(t):
        lo, hi = int(bounds[t]), int(bounds[t + 1])
        if hi <= lo:
            return
        a = list(args)
        a[0] = hi - lo
        for p, k in enumerate(sig):
            if (k == "P" and p < first_j) or k == "O":
                a[p] = args[p][lo:hi]
        call(lib, name, prec, *a)

    with ThreadPoolExecutor(nthreads) as ex:
        list(ex.map(piece, range(nthreads)))
