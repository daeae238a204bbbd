"""Wall time of one call through the reference-facing seam (host numpy arrays in and out,
synchronous: ParticleSystem.set_acc_jerk -> extensions.AccJerk.calc -> CUDAKernel -> C ABI)
for small and medium N, with the library's own stage times (H2D, pack, pair, finalize, D2H).
This is what the unmodified reference integrators pay per force evaluation.

    python tools/latency_probe.py [kernel ...]
"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from tupan_b200 import backend, ics  # noqa: E402


def main():
    lib = backend.require_gpu("float64")
    for n in (64, 256, 1024, 4096, 16384, 65536):
        ps = ics.make_plummer(n, seed=1)
        for name, call in (("acc_jerk", lambda: ps.set_acc_jerk(ps)),
                           ("tstep", lambda: ps.set_tstep(ps, 1.0 / 64)),
                           ("phi", lambda: ps.set_phi(ps))):
            for _ in range(5):
                call()
            reps = 200 if n <= 4096 else 20
            t0 = time.perf_counter()
            for _ in range(reps):
                call()
            wall = (time.perf_counter() - t0) / reps
            lib.tupan_cuda_set_timing(1)
            call()
            st = [ctypes.c_float() for _ in range(5)]
            lib.tupan_cuda_last_times(*[ctypes.byref(x) for x in st])
            lib.tupan_cuda_set_timing(0)
            print("N=%-6d %-9s wall %8.1f us/call   stages(us): h2d %.1f pack %.1f pair %.1f finalize %.1f d2h %.1f"
                  % ((n, name, wall * 1e6) + tuple(x.value * 1e3 for x in st)), flush=True)


if __name__ == "__main__":
    main()
