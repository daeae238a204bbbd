"""BASELINE.json configs[3]: acc_jerk fp64 throughput sweep, N = 2^14 ... 2^22, on 1/2/4/8 GPUs.

    python tools/sweep.py [lo=14] [hi=22] [prec=float64] [kernel=acc_jerk_kernel]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port 29577 tools/sweep.py ...

Device-resident state, CUDA events on the launching stream, max over ranks, >= 3 warm-up
evaluations at small N (1 at the largest), L2 flushed between evaluations.  Prints one line
per N: ms, Gpair/s (all ranks), TFLOP/s under the reference's flop convention and the
fraction of the FMA-pipe peak measured in the same process."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tupan_b200 import device, ics, sharded  # noqa: E402

FLOPS = {"phi_kernel": 14, "acc_kernel": 20, "acc_jerk_kernel": 42, "snap_crackle_kernel": 114,
         "tstep_kernel": 42, "nreg_Xkernel": 37, "nreg_Vkernel": 25}


def main():
    ap = argparse.ArgumentParser()
    # positional on purpose: torchrun's own parser abbreviates-matches long options that follow
    # the script name (--lo -> --log-dir ...)
    ap.add_argument("lo", type=int, nargs="?", default=14)
    ap.add_argument("hi", type=int, nargs="?", default=22)
    ap.add_argument("prec", nargs="?", default="float64")
    ap.add_argument("kernel", nargs="?", default="acc_jerk_kernel")
    ap.add_argument("transport", nargs="?", default="auto", help="auto | nccl | p2p | p2p-graph")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float64 if args.prec == "float64" else torch.float32
    peak, mhz = device.fma_peak(args.prec, 300.0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    scal = (1.0 / 64,) if args.kernel in ("tstep_kernel", "nreg_Xkernel", "nreg_Vkernel") else ()
    if rank == 0:
        print("# %s %s on %d x %s, transport %s; FMA-pipe (3-register DFMA chain) probe in-process: %.2f TFLOP/s per "
              "GPU (%.0f MHz effective)" % (args.kernel, args.prec, world, torch.cuda.get_device_name(dev),
                                            args.transport, peak, mhz))
        print("# %8s %10s %12s %10s %8s %8s" % ("N", "ms", "Gpair/s", "TFLOP/s", "%peak", "per-GPU"))
    for p in range(args.lo, args.hi + 1):
        n = 1 << p
        ps = ics.make_plummer(n, seed=1, dtype=args.prec)
        full = device.to_device(ps, device=dev)
        for k in ("ax", "ay", "az", "jx", "jy", "jz"):
            full[k] = torch.zeros(n, dtype=dtype, device=dev)
        if world > 1:
            sk = sharded.ShardedKernel(args.kernel, n, dtype, dev, transport=args.transport.split("-")[0])
            mine = {a: full[a][sk.lo:sk.hi].contiguous() for a in device.KERNEL_INPUTS[args.kernel]}
            out = sk.evaluate(mine, scal)
            graphed = args.transport.endswith("-graph") and sk.transport == "p2p"

            def step():
                if graphed:
                    sk.evaluate_graphed(mine, scal, out)
                else:
                    sk.evaluate(mine, scal, out)
        else:
            out = device.run(args.kernel, full, full, scal)

            def step():
                device.run(args.kernel, full, full, scal, out)
        est_ms = float(n) * n / 490e9 * 1e3 / world
        warm = 3 if est_ms < 3000 else 1
        reps = max(2, min(20, int(2000 / max(est_ms, 0.05))))
        for _ in range(warm):
            step()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for e0, e1 in ev:
            flush.zero_()
            e0.record()
            step()
            e1.record()
        torch.cuda.synchronize()
        ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev)[len(ev) // 2]          # median
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        if rank == 0:
            pairs = float(n) * n / (ms * 1e-3)
            tf = pairs * FLOPS[args.kernel] * 1e-12
            print("  %8d %10.3f %12.2f %10.2f %7.1f%% %8.2f  %s" % (n, ms, pairs * 1e-9, tf, 100 * tf / (peak * world),
                                                                    pairs * 1e-9 / world,
                                                                    sk.transport if world > 1 else "-"), flush=True)
        if world > 1 and sk.transport == "p2p":
            sk.check()
        del full, out
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
