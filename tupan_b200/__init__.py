"""tupan_b200 -- B200-native (sm_100a) pairwise gravity kernels behind tupan's kernel interface.

Scope: the O(N^2) kernels of ggf84/tupan's ``tupan/lib`` (phi, acc, acc_jerk, snap_crackle,
tstep, pnacc, nreg_X/V, sakura, kepler) as hand-written CUDA, exposed through the reference's
own C ABI (``include/libtupan_cuda.h``) and kernel-adapter protocol.  See DESIGN.md.
"""
__version__ = "0.1.0"

from . import backend, extensions, ics, particles  # noqa: F401
from .backend import CUDAKernel, TupanCudaError  # noqa: F401
from .particles import ParticleSystem  # noqa: F401
