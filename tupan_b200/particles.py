"""SoA particle container with the attribute names the reference kernels are called with.

Mirrors the part of ``tupan/particles`` that the kernel wrappers touch
(``particles/body.py:26-39`` attribute list, ``allparticles.py:64-76``
``register_auxiliary_attribute``, ``body.py:324-361`` ``set_*`` force setters): one 1-D numpy
array per attribute, outputs created lazily on the i-system.  It exists so that the
wrappers in :mod:`tupan_b200.extensions` can be driven without the reference installed
(the GPU box has no ``/root/reference``); an unmodified ``tupan.particles.ParticleSystem``
works with the same wrappers because only attribute access is used.
"""
import copy

import numpy as np

BASE_ATTRS = ("id", "mass", "eps2", "rx", "ry", "rz", "vx", "vy", "vz",
              "time", "nstep", "tstep")
_INT_ATTRS = {"id": "uint", "nstep": "uint"}


class ParticleSystem(object):
    def __init__(self, n=0, dtype=np.float64):
        self.n = int(n)
        self.dtype = np.dtype(dtype)
        utype = np.uint64 if self.dtype == np.float64 else np.uint32
        for name in BASE_ATTRS:
            dt = utype if name in _INT_ATTRS else self.dtype
            setattr(self, name, np.zeros(self.n, dt))
        self.id[...] = np.arange(self.n)

    # -- reference API: allparticles.py:64-76 ------------------------------------------
    def register_auxiliary_attribute(self, attr, sctype):
        if attr in self.__dict__:
            raise ValueError("'{0}' is already a registered attribute.".format(attr))
        utype = np.uint64 if self.dtype == np.float64 else np.uint32
        itype = np.int64 if self.dtype == np.float64 else np.int32
        dt = {"real": self.dtype, "uint": utype, "int": itype}[sctype]
        setattr(self, attr, np.zeros(self.n, dt))

    def arrays(self):
        return {k: v for k, v in self.__dict__.items()
                if isinstance(v, np.ndarray) and v.shape == (self.n,)}

    def __len__(self):
        return self.n

    def copy(self):
        return copy.deepcopy(self)

    def __getitem__(self, slc):
        if isinstance(slc, (int, np.integer)):
            slc = slice(slc, slc + 1) if slc != -1 else slice(slc, None)
        out = type(self).__new__(type(self))
        out.dtype = self.dtype
        for k, v in self.arrays().items():
            setattr(out, k, np.ascontiguousarray(v[slc]))
        out.n = len(out.mass)
        return out

    def astype(self, dtype):
        out = type(self)(self.n, dtype)
        for k, v in self.arrays().items():
            if v.dtype.kind == "f":
                setattr(out, k, v.astype(out.dtype))
        return out

    # -- reference API: body.py:324-361 ------------------------------------------------
    def set_tstep(self, ps, eta):
        from . import extensions
        extensions.get("tstep", self.dtype).calc(self, ps, eta)

    def set_phi(self, ps):
        from . import extensions
        extensions.get("phi", self.dtype).calc(self, ps)

    def set_acc(self, ps):
        from . import extensions
        extensions.get("acc", self.dtype).calc(self, ps)

    def set_pnacc(self, ps):
        from . import extensions
        extensions.get("pnacc", self.dtype).calc(self, ps)

    def set_acc_jerk(self, ps):
        from . import extensions
        extensions.get("acc_jerk", self.dtype).calc(self, ps)

    def set_snap_crackle(self, ps):
        from . import extensions
        extensions.get("snap_crackle", self.dtype).calc(self, ps)

    # -- diagnostics used by the energy-error checks (body.py:262-306) -------------------
    @property
    def kinetic_energy(self):
        v2 = self.vx ** 2 + self.vy ** 2 + self.vz ** 2
        return float(0.5 * np.sum(self.mass.astype(np.float64) * v2))

    @property
    def potential_energy(self):
        self.set_phi(self)
        return float(0.5 * np.sum(self.mass.astype(np.float64) * self.phi))
