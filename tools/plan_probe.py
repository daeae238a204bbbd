"""Launch-plan sweep for small and medium N (GPU box): device-resident time of one kernel under
every forced plan (throughput shape with jg chunks; lane split with 2^js lanes per particle and
jg chunks) next to the plan choose_plan() picks.  Feeds the cost model in pair_engine.cuh.

    python tools/plan_probe.py [kernel] [prec] [N,N,...]
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tupan_b200 import backend, device, ics  # noqa: E402


def timed(kern, d, scal, out, reps):
    device.run(kern, d, d, scal, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        device.run(kern, d, d, scal, out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    kern = sys.argv[1] if len(sys.argv) > 1 else "acc_jerk_kernel"
    prec = sys.argv[2] if len(sys.argv) > 2 else "float64"
    sizes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [512, 1024, 2048, 4096, 8192, 16384]
    scal = {"tstep_kernel": (1 / 64,)}.get(kern, ())
    lib = backend.require_gpu(prec)
    for n in sizes:
        ps = ics.make_plummer(n, seed=1, dtype=prec)
        d = device.to_device(ps)
        out = device.run(kern, d, d, scal)
        reps = 50 if n <= 4096 else 10
        lib.tupan_cuda_force_plan(-1, 0, 1)
        t = timed(kern, d, scal, out, reps)
        pl = [ctypes.c_int() for _ in range(3)]
        lib.tupan_cuda_last_plan(*[ctypes.byref(x) for x in pl])
        print("N=%-6d %s chosen plan (split=%d js=%d jg=%d): %8.1f us  %7.1f Gpair/s"
              % (n, kern, pl[0].value, pl[1].value, pl[2].value, t, float(n) * n / t * 1e-3), flush=True)
        rows = []
        tiles = (n + 127) // 128
        for jg in (1, 2, 4, 8, 16, 32, 64):
            if jg > tiles:
                continue
            lib.tupan_cuda_force_plan(0, 0, jg)
            rows.append(("thr jg=%d" % jg, timed(kern, d, scal, out, reps)))
        for js in (0, 1, 2, 3, 4, 5):
            for jg in (1, 2, 4, 8, 16, 32):
                if jg > tiles:
                    continue
                lib.tupan_cuda_force_plan(1, js, jg)
                rows.append(("split js=%d jg=%d" % (js, jg), timed(kern, d, scal, out, reps)))
        lib.tupan_cuda_force_plan(-1, 0, 1)
        rows.sort(key=lambda r: r[1])
        print("   best: " + "; ".join("%s %.1f us" % r for r in rows[:6]), flush=True)
        print("   all : " + "; ".join("%s %.1f" % r for r in sorted(rows)), flush=True)


if __name__ == "__main__":
    main()
