"""Kernel wrappers with the call interface of the reference's ``tupan/lib/extensions.py``.

Same class names (``Phi, Acc, AccJerk, SnapCrackle, Tstep, PNAcc, Sakura, NREG_X, NREG_V,
Kepler``), same ``calc(ips, jps, *scalars)`` = ``set_args -> run -> get_result`` sequence
(``extensions.py:84-97``), same argument marshalling (which particle arrays, in which order)
and the same lazily registered output attributes on ``ips`` (``ax..az, jx..jz, sx..sz,
cx..cz, phi, tstep, tstepij, pnax.., drx.., dvx.., mrx.., u, mvx.., mk``).  The only backend
is CUDA.  ``install(tupan.lib.extensions)`` swaps these kernels into an imported reference
tree so that its unmodified integrators drive the GPU path.
"""
import numpy as np

from .backend import CUDAKernel, prec_of

__all__ = ["Phi", "Acc", "AccJerk", "SnapCrackle", "Tstep", "PNAcc", "Sakura", "NREG_X", "NREG_V",
           "Kepler", "Clight", "clight", "get", "install"]


class Clight(object):
    """PN order, speed of light and its inverse powers (reference: extensions.py:31-60)."""

    def __init__(self):
        self._pn_order = 0
        self._clight = None
        self.inv1 = self.inv2 = self.inv3 = self.inv4 = self.inv5 = self.inv6 = self.inv7 = 0.0

    @property
    def pn_order(self):
        return self._pn_order

    @pn_order.setter
    def pn_order(self, value):
        self._pn_order = int(value)

    @property
    def clight(self):
        return self._clight

    @clight.setter
    def clight(self, value):
        self._clight = float(value)
        self.inv1 = 1.0 / self._clight
        self.inv2 = self.inv1 ** 2
        self.inv3 = self.inv1 ** 3
        self.inv4 = self.inv1 ** 4
        self.inv5 = self.inv1 ** 5
        self.inv6 = self.inv1 ** 6
        self.inv7 = self.inv1 ** 7


clight = Clight()

_STATE8 = ("mass", "rx", "ry", "rz", "eps2", "vx", "vy", "vz")
_STATE5 = ("mass", "rx", "ry", "rz", "eps2")


class _Extension(object):
    """One kernel: `name` in libtupan.h, particle attributes read on each side, scalar
    converters, and output attributes registered on the i-system."""
    kernel_name = None
    inputs = ()
    outputs = ()
    scalars = ()          # names of cty converters, e.g. ("c_real",)
    uses_gsize = True

    def __init__(self, backend="CUDA", prec="float64"):
        if backend != "CUDA":
            raise ValueError("Inappropriate 'backend': {}. Supported values: ['CUDA']".format(backend))
        self.prec = prec_of(prec)
        self.kernel = CUDAKernel(self.prec, self.kernel_name)
        cty = self.kernel.cty
        n = len(self.inputs)
        self.argtypes = ((cty.c_uint,) + (cty.c_real_p,) * n) * 2 + tuple(getattr(cty, s) for s in self.scalars)
        self.restypes = (cty.c_real_p,) * len(self.outputs)

    def _scalars(self, *args):
        return args

    def set_args(self, ips, jps, *args):
        ni, nj = ips.n, jps.n
        if self.uses_gsize:
            self.kernel.set_gsize(ni, nj)
        for attr in self.outputs:
            if attr not in ips.__dict__:
                ips.register_auxiliary_attribute(attr, "real")
        self._inargs = ((ni,) + tuple(getattr(ips, a) for a in self.inputs)
                        + (nj,) + tuple(getattr(jps, a) for a in self.inputs)
                        + tuple(self._scalars(*args)))
        self._outargs = tuple(getattr(ips, a) for a in self.outputs)
        self.inargs = [t(a) for a, t in zip(self._inargs, self.argtypes)]
        self.outargs = [t(a) for a, t in zip(self._outargs, self.restypes)]
        self.kernel.set_args(self.inargs + self.outargs)

    def run(self):
        self.kernel.run()

    def get_result(self):
        return self.kernel.map_buffers(self._outargs, self.outargs)

    def calc(self, ips, jps, *args):
        self.set_args(ips, jps, *args)
        self.run()
        return self.get_result()


class Phi(_Extension):            # reference: extensions.py:100-160
    kernel_name = "phi_kernel"
    inputs = _STATE5
    outputs = ("phi",)


class Acc(_Extension):            # :163-236
    kernel_name = "acc_kernel"
    inputs = _STATE5
    outputs = ("ax", "ay", "az")


class AccJerk(_Extension):        # :239-285
    kernel_name = "acc_jerk_kernel"
    inputs = _STATE8
    outputs = ("ax", "ay", "az", "jx", "jy", "jz")


class SnapCrackle(_Extension):    # :288-346
    kernel_name = "snap_crackle_kernel"
    inputs = _STATE8 + ("ax", "ay", "az", "jx", "jy", "jz")
    outputs = ("sx", "sy", "sz", "cx", "cy", "cz")


class Tstep(_Extension):          # :349-392  calc(ips, jps, eta)
    kernel_name = "tstep_kernel"
    inputs = _STATE8
    outputs = ("tstep", "tstepij")
    scalars = ("c_real",)


class PNAcc(_Extension):          # :395-446  scalars come from the module-level `clight`
    kernel_name = "pnacc_kernel"
    inputs = _STATE8
    outputs = ("pnax", "pnay", "pnaz")
    scalars = ("c_uint",) + ("c_real",) * 7

    def _scalars(self):
        c = clight
        return (c.pn_order, c.inv1, c.inv2, c.inv3, c.inv4, c.inv5, c.inv6, c.inv7)


class Sakura(_Extension):         # :449-511  calc(ips, jps, dt, flag)
    kernel_name = "sakura_kernel"
    inputs = _STATE8
    outputs = ("drx", "dry", "drz", "dvx", "dvy", "dvz")
    scalars = ("c_real", "c_int")
    uses_gsize = False


class NREG_X(_Extension):         # :514-570  calc(ips, jps, dt)
    kernel_name = "nreg_Xkernel"
    inputs = _STATE8
    outputs = ("mrx", "mry", "mrz", "ax", "ay", "az", "u")
    scalars = ("c_real",)


class NREG_V(_Extension):         # :573-617  calc(ips, jps, dt)
    kernel_name = "nreg_Vkernel"
    inputs = ("mass", "vx", "vy", "vz", "ax", "ay", "az")
    outputs = ("mvx", "mvy", "mvz", "mk")
    scalars = ("c_real",)


class Kepler(_Extension):         # :620-651  calc(ips, jps, dt); two bodies, in place
    kernel_name = "kepler_solver_kernel"
    inputs = _STATE8
    outputs = ("rx", "ry", "rz", "vx", "vy", "vz")
    scalars = ("c_real",)

    def __init__(self, backend="CUDA", prec="float64"):
        _Extension.__init__(self, backend, prec)
        cty = self.kernel.cty
        self.argtypes = (cty.c_real_p,) * 8 + (cty.c_real,)

    def set_args(self, ips, jps, dt):
        if ips.n != 2:
            raise ValueError("kepler_solver_kernel propagates exactly two bodies, got %d" % ips.n)
        self._inargs = tuple(getattr(ips, a) for a in self.inputs) + (dt,)
        self._outargs = tuple(getattr(ips, a) for a in self.outputs)   # outputs alias inputs
        self.inargs = [t(a) for a, t in zip(self._inargs, self.argtypes)]
        self.outargs = [t(a) for a, t in zip(self._outargs, self.restypes)]
        self.kernel.set_args(self.inargs + self.outargs)


_CLASSES = {"phi": Phi, "acc": Acc, "acc_jerk": AccJerk, "snap_crackle": SnapCrackle, "tstep": Tstep,
            "pnacc": PNAcc, "sakura": Sakura, "nreg_x": NREG_X, "nreg_v": NREG_V, "kepler": Kepler}
_singletons = {}


def get(name, dtype=np.float64):
    """Module-level kernel objects, created on first use (the reference creates its own at
    import time, extensions.py:654-666)."""
    key = (name, prec_of(dtype))
    if key not in _singletons:
        _singletons[key] = _CLASSES[name]("CUDA", key[1])
    return _singletons[key]


def install(ref_extensions, prec=None):
    """Swap the CUDA kernels into an imported reference ``tupan.lib.extensions`` module, so
    that ``tupan.particles`` / ``tupan.integrator`` run unchanged on the GPU path.  The
    reference's own ``clight`` object is kept and mirrored (its PNAcc reads it at call time)."""
    if prec is None:
        from importlib import import_module
        prec = import_module(ref_extensions.__name__.rsplit(".", 1)[0] + ".utils.ctype").prec
    global clight
    clight = ref_extensions.clight
    for name in _CLASSES:
        setattr(ref_extensions, name, get(name, prec))
    ref_extensions.backend = "CUDA"
    return ref_extensions
