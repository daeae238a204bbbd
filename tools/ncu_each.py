"""One launch of every pair kernel (both precisions) between cudaProfilerStart/Stop, for
    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:pair_kernel -o gpurun_out/r02_all_kernels python tools/ncu_each.py [N]
so that one capture holds one profiled launch per kernel.  Summarise with tools/ncu_summary.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tupan_b200 import device, ics  # noqa: E402

KERNELS = ["phi_kernel", "acc_kernel", "acc_jerk_kernel", "snap_crackle_kernel", "tstep_kernel", "nreg_Xkernel",
           "nreg_Vkernel", "pnacc_kernel", "sakura_kernel"]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    for prec in ("float64", "float32"):
        ps = ics.make_plummer(n, seed=1, dtype=prec)
        d = device.to_device(ps)
        for k in ("ax", "ay", "az", "jx", "jy", "jz"):
            d[k] = torch.randn(n, dtype=d["mass"].dtype, device="cuda")
        for kern in KERNELS:
            scal = {"tstep_kernel": (1 / 64,), "nreg_Xkernel": (1 / 64,), "nreg_Vkernel": (1 / 64,),
                    "sakura_kernel": (1 / 1024, 1),
                    "pnacc_kernel": (7,) + tuple(128.0 ** -k for k in range(1, 8))}.get(kern, ())
            out = device.run(kern, d, d, scal)            # warm-up, buffers sized
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            device.run(kern, d, d, scal, out)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
