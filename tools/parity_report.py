"""Print the measured CUDA-vs-oracle deviation of every kernel (run on the GPU box).
Used to set / re-check the tolerances stated in tests/test_parity_gpu.py."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import oracle  # noqa: E402
from tupan_b200 import backend, ics  # noqa: E402
from util import KERNELS, as_dict, cuda_run, pn_scalars, rel_err, run  # noqa: E402


def main():
    rows = []
    for prec in ("float64", "float32"):
        lib = backend.require_gpu(prec)
        olib = oracle.load("oracle", prec)
        sets = {
            "uniform256_eps0": as_dict(ics.make_uniform(256, seed=1), prec),
            "plummer1024": as_dict(ics.make_plummer(1024, seed=1), prec),
            "plummer4096": as_dict(ics.make_plummer(4096, seed=2), prec),
        }
        for sname, data in sets.items():
            n = len(data["mass"])
            for name in KERNELS:
                variants = [None]
                if name == "pnacc_kernel":
                    variants = [pn_scalars(k) for k in (2, 4, 5, 6, 7)]
                if name == "sakura_kernel":
                    variants = [(1.0 / 64, f) for f in (-2, -1, 1, 2)]
                    if n > 1024:
                        continue
                for sc in variants:
                    ref = run(olib, name, prec, data, data, sc)
                    for plan in ((-1, 0, 1), (0, 0, 1), (1, 0, 1), (1, 3, 1), (1, 5, 3), (0, 0, 2)):
                        lib.tupan_cuda_force_plan(*plan)
                        t0 = time.time()
                        got = cuda_run(name, prec, data, data, sc)
                        dt = time.time() - t0
                        e = rel_err(name, got, ref)
                        rows.append((prec, sname, name, sc[0] if sc else "", plan, e, dt))
                        print("%-8s %-16s %-20s %-8s plan=%-12s err=%.3e  %.1f ms" % (
                            prec, sname, name, (sc[:2] if sc else ""), plan, e, dt * 1e3), flush=True)
                    lib.tupan_cuda_force_plan(-1, 0, 1)
    worst = {}
    for prec, sname, name, v, plan, e, dt in rows:
        worst[(prec, name)] = max(worst.get((prec, name), 0.0), e)
    print("\nworst per kernel:")
    for k, v in sorted(worst.items()):
        print("  %-8s %-22s %.3e" % (k[0], k[1], v))


if __name__ == "__main__":
    main()
