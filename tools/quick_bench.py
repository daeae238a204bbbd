"""Quick device-resident throughput probe of every kernel (GPU box)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tupan_b200 import backend, device, ics  # noqa: E402

FLOPS = {"phi_kernel": 14, "acc_kernel": 20, "acc_jerk_kernel": 42, "snap_crackle_kernel": 114,
         "tstep_kernel": 42, "pnacc_kernel": 632, "nreg_Xkernel": 37, "nreg_Vkernel": 25, "sakura_kernel": 0}


def main():
    sizes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "16384,65536,262144".split(","))]
    kernels = sys.argv[2].split(",") if len(sys.argv) > 2 else ["acc_jerk_kernel"]
    precs = sys.argv[3].split(",") if len(sys.argv) > 3 else ["float64"]
    for prec in precs:
        tf, mhz = device.fma_peak(prec)
        print("%s FMA peak: %.2f TFLOP/s (effective SM clock %.0f MHz)" % (prec, tf, mhz), flush=True)
        for n in sizes:
            ps = ics.make_plummer(n, seed=1, dtype=prec)
            d = device.to_device(ps)
            for k in ("ax", "ay", "az", "jx", "jy", "jz"):
                d[k] = torch.randn(n, dtype=d["mass"].dtype, device="cuda")
            for kern in kernels:
                scal = {"tstep_kernel": (1 / 64,), "nreg_Xkernel": (1 / 64,), "nreg_Vkernel": (1 / 64,),
                        "sakura_kernel": (1 / 64, 1),
                        "pnacc_kernel": (7,) + tuple(128.0 ** -k for k in range(1, 8))}.get(kern, ())
                out = device.run(kern, d, d, scal)
                torch.cuda.synchronize()
                reps = 3
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    device.run(kern, d, d, scal, out)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                pairs = float(n) * n / (ms * 1e-3)
                print("%-8s %-20s N=%-8d %9.3f ms  %8.2f Gpair/s  %6.2f TFLOP/s (%4.1f%% of FMA peak)" % (
                    prec, kern, n, ms, pairs * 1e-9, pairs * FLOPS[kern] * 1e-12,
                    100 * pairs * FLOPS[kern] * 1e-12 / tf), flush=True)


if __name__ == "__main__":
    main()
