#!/usr/bin/env python
"""Headline benchmark: acc_jerk pair-interactions/s, fp64, Plummer N = 2^20 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference] [--n N]

One "step" = one acc_jerk evaluation of all N x N pairs (masked self pairs included, as the
reference loop evaluates them) -- one force evaluation of a Hermite step.

* ``value``    pairs/s, particle state resident in HBM (device-pointer API), CUDA events,
               K steps back to back, max over ranks; L2 flushed between steps.
* ``e2e``      the same metric through the reference-facing path: ``ps.set_acc_jerk(ps)`` ->
               ``extensions.AccJerk.calc`` -> ``CUDAKernel`` -> C ABI ``acc_jerk_kernel`` with
               HOST (pinned) numpy arrays; H2D of the inputs and D2H of the six outputs happen
               inside every timed step.
* ``roofline`` FP64 FMA pipe (this path is FMA-bound, not HBM- or tensor-bound; the bytes are
               reported for completeness): 42 flop/pair (the reference's own convention,
               acc_jerk_kernel_common.h:56) x pairs / pair-kernel time, against the FP64 FMA
               rate this board sustains in a pure-DFMA probe run in the same process.
* ``cpu_baseline`` the reference's C backend (oracle/_ref, else the oracle port) on the host
               cores, on a bounded i-sample against the full j-set.
* ``parity``   after the timed region (never inside it) every rank checks the outputs of its LAST
               timed launch -- and of the last end-to-end call -- on an i-sample of its shard
               against the reference C backend with the full j-set: per-particle norm-relative
               error of acc and of jerk, maximum over the sample and over ranks; tolerance 1e-12.

With N > 1 (torchrun, one rank per GPU) the i-set is sharded, the packed j rows are
all-gathered over NCCL once per step, total work is fixed: "scaling": "strong".
``--impl reference`` times the reference's CPU implementation only (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FLOPS_PER_PAIR = 42          # acc_jerk_kernel_common.h:56 "Total flop count: 42"
DP_INSTR_PER_PAIR = 31       # FP64-pipe instructions our kernel issues per pair (SASS count)
S8 = ("mass", "rx", "ry", "rz", "eps2", "vx", "vy", "vz")
OUT6 = ("ax", "ay", "az", "jx", "jy", "jz")
METRIC = "acc_jerk pair-interactions/s fp64"
# dram__bytes_read.sum + dram__bytes_write.sum of the pair kernel, per launch, from ncu captures of
# this command, keyed by (n, GPUs, j-chunks of the launch):
#   round 2, grouped 3x2 kernel, 17 chunks (profiles/r02_accjerk_grouped_n1m_dram.csv): 1.082 GB + 0.877 GB
#   round 1, 25 chunks (profiles/r01_accjerk_v4_n1m_ncu_summary.txt): 1.568 GB + 1.358 GB
# It exceeds the 184.5 MB of algorithmic bytes because the j range is split into chunks for wave
# balance (each chunk writes a partial accumulator slot that finalize reads back: 17 x 2^20 x 6 x 8 B =
# 0.855 GB); at 1 GB/s it is 0.02 % of HBM bandwidth -- this kernel is FP64-pipe bound.
NCU_TRAFFIC = {(1 << 20, 1, 17): 1082319872 + 877085440, (1 << 20, 1, 25): 1567549000 + 1357939000}
KERNEL_NAME = "pair_kernel_grouped<AccJerkOp<double>>"
PARITY_TOL = 1e-12           # BASELINE.json north_star: ~1e-12 in fp64 (summation order differs)
PARITY_SAMPLE = 256          # i-particles per rank checked against the full j-set


def workload_config(n):
    """The part of `config` both arms print identically."""
    return {"workload": "Plummer sphere N=%d equal-mass, eps=4/N, acc_jerk fp64 (one Hermite force "
                        "evaluation)" % n, "n": n}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=("cuda", "reference"))
    ap.add_argument("--n", "--particles", dest="n", type=int, default=1 << 20,
                    help="particles (BASELINE: 2^20); spell it --particles under torchrun, whose own parser\n"
                         "takes --n for an abbreviation of its options")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--transport", default=None, choices=("nccl", "p2p"),
                    help="multi-GPU j rows: NCCL all-gather (default) or read in place through peer mappings")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ---------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[6]))
            except ValueError:
                continue
            for name, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


# ---------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's C backend on the host cores
# ---------------------------------------------------------------------------------------
def cpu_reference_lib():
    import oracle
    if oracle.have("ref", "float64"):
        return oracle.load("ref", "float64"), "reference"
    return oracle.load("oracle", "float64"), "port"


def cpu_sample(ps, n, lib, cores, ni_sample):
    """acc_jerk of `ni_sample` i-particles against all n j-particles, `cores` threads (the
    same unmodified C function on contiguous i-slices).  Returns (seconds, pairs)."""
    import oracle
    idx = np.linspace(0, n - 1, ni_sample).astype(np.int64)
    ia = [np.ascontiguousarray(getattr(ps, a)[idx]) for a in S8]
    ja = [getattr(ps, a) for a in S8]
    outs = [np.zeros(ni_sample) for _ in range(6)]
    args = [ni_sample] + ia + [n] + ja + outs
    t0 = time.perf_counter()
    oracle.call_threaded(lib, "acc_jerk_kernel", "float64", cores, *args)
    return time.perf_counter() - t0, float(ni_sample) * n


def cpu_sample_size(n, cores, seconds=4.0, rate_per_core=0.9e8):
    ni = int(seconds * rate_per_core * cores / n)
    return max(cores, min(n, (ni // cores) * cores))


def parity_sample(ps, n, lo, hi, got, lib, cores, nsample=PARITY_SAMPLE):
    """max over a sample of this shard's particles of ||got - ref|| / ||ref|| (acc and jerk
    separately), ref = the reference C backend on the same inputs with the full j-set.
    `got` = six arrays holding the shard's outputs (index 0 = particle lo)."""
    import oracle
    ni = hi - lo
    if ni <= 0:
        return 0.0, 0
    k = min(nsample, ni)
    rng = np.random.default_rng(1234 + lo)
    idx = np.unique(np.concatenate([np.linspace(0, ni - 1, k // 2).astype(np.int64),
                                    rng.integers(0, ni, k - k // 2)]))
    ia = [np.ascontiguousarray(getattr(ps, a)[lo + idx]) for a in S8]
    ja = [np.ascontiguousarray(getattr(ps, a)) for a in S8]
    ref = [np.zeros(len(idx)) for _ in range(6)]
    oracle.call_threaded(lib, "acc_jerk_kernel", "float64", max(1, min(cores, len(idx))),
                         *([len(idx)] + ia + [n] + ja + ref))
    worst = 0.0
    for g0 in (0, 3):
        g = np.stack([np.asarray(got[g0 + c])[idx] for c in range(3)])
        r = np.stack(ref[g0:g0 + 3])
        if not np.all(np.isfinite(g)):
            return float("inf"), len(idx)
        worst = max(worst, float(np.max(np.sqrt(((g - r) ** 2).sum(0)) / np.sqrt((r ** 2).sum(0)))))
    return worst, len(idx)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tupan_b200 import ics
    n = args.n
    ps = ics.make_plummer(n, seed=1)
    lib, kind = cpu_reference_lib()
    cores = os.cpu_count() or 1
    ni = cpu_sample_size(n, cores)
    cpu_sample(ps, n, lib, cores, ni)            # one warm-up sample (page-in, thread pool)
    t = 0.0
    pairs = 0.0
    for _ in range(args.steps):
        dt, p = cpu_sample(ps, n, lib, cores, ni)
        t += dt
        pairs += p
    v = pairs / t
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n),
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": kind,
                         "sample": "each step = ni=%d i-particles (evenly spaced) x nj=%d, %d threads on "
                                   "contiguous i-slices of the unmodified C function" % (ni, n, cores)},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------
def pinned_system(ps, torch):
    """Copy the particle arrays (and pre-register the outputs) into pinned host memory."""
    keep = []
    for a in S8 + OUT6:
        src = getattr(ps, a, None)
        t = torch.empty(ps.n, dtype=torch.float64).pin_memory()
        if src is not None:
            t.copy_(torch.from_numpy(src))
        else:
            t.zero_()
        keep.append(t)
        setattr(ps, a, t.numpy())
    ps._pinned = keep
    return ps


def run_cuda(args):
    import torch
    import torch.distributed as dist
    from tupan_b200 import backend, device, ics, sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # convenience: relaunch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                   "--master-addr", "127.0.0.1", "--master-port", "29533"] + sys.argv
            sys.exit(subprocess.call(cmd))
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = backend.require_gpu("float64")       # raises if the CUDA library is missing / no GPU

    n = args.n
    ps = ics.make_plummer(n, seed=1)
    bounds = sharded.shard_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    ni = hi - lo

    # ---- device-resident state ------------------------------------------------------------
    full = device.to_device({a: getattr(ps, a) for a in S8}, device=dev)
    local = {a: full[a][lo:hi].contiguous() for a in S8}
    out = {a: torch.empty(ni, dtype=torch.float64, device=dev) for a in OUT6}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    if world > 1:
        sk = sharded.ShardedKernel("acc_jerk_kernel", n, torch.float64, dev, transport=args.transport)

        def step():
            flush.zero_()
            sk.evaluate(local, (), out)
    else:
        def step():
            flush.zero_()
            device.run("acc_jerk_kernel", full, full, (), out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak_tf, peak_mhz = device.fma_peak("float64", 300.0)

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = lib.tupan_cuda_launch_count()
    lib.tupan_cuda_set_timing(1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1 and sk.transport == "p2p":
        sk.peer.check()                      # a peer barrier that gave up waiting voids the run: fail loudly
    ms = e0.elapsed_time(e1)
    launches = lib.tupan_cuda_launch_count() - launches0
    # pair-kernel time: the library records CUDA events around every pair-kernel launch on the
    # launching stream; summed over ALL launches of the timed region, divided by the steps
    sums = (ctypes.c_float * 5)()
    calls = ctypes.c_longlong()
    wrapped = lib.tupan_cuda_sum_times(sums, ctypes.byref(calls))
    pair_ms = sums[2] / args.steps if (calls.value > 0 and not wrapped) else 0.0
    pair_launches_per_step = calls.value / float(args.steps)
    lib.tupan_cuda_set_timing(0)
    plan = [ctypes.c_int() for _ in range(3)]
    lib.tupan_cuda_last_plan(*[ctypes.byref(x) for x in plan])
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = float(n) * n / (ms_per_step * 1e-3)

    # ---- parity of the outputs of the last timed launch (outside the timed region) -------------
    plib, pkind = cpu_reference_lib()
    cores = os.cpu_count() or 1
    got = [out[a].cpu().numpy() for a in OUT6]
    perr, pn = parity_sample(ps, n, lo, hi, got, plib, max(1, cores // max(1, min(world, 8))))
    perr_e2e = None

    # ---- end to end through the reference-facing API with host arrays --------------------------
    e2e = None
    if not args.no_e2e:
        hps = pinned_system(ps, torch)
        ips = hps if world == 1 else hps[lo:hi]
        for a in OUT6:                       # outputs of the shard: views of the pinned arrays
            setattr(ips, a, getattr(hps, a)[lo:hi])
        ips.set_acc_jerk(hps)                # warm-up (buffers sized)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ips.set_acc_jerk(hps)            # H2D i + j, pack, pair kernel, D2H, synchronous
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        perr_e2e, _ = parity_sample(ps, n, lo, hi, [getattr(ips, a) for a in OUT6], plib,
                                    max(1, cores // max(1, min(world, 8))))
        h2d = 8 * ni * 8 + (0 if world == 1 else 8 * n * 8)   # j arrays alias the i arrays at N=1
        e2e = {"value": float(n) * n / (dt / args.steps), "unit": "pairs/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(6 * ni * 8),
               "api": "ParticleSystem.set_acc_jerk -> extensions.AccJerk.calc -> CUDAKernel -> acc_jerk_kernel "
                      "(C ABI, pinned host arrays)"}

    pt = torch.tensor([perr, perr_e2e if perr_e2e is not None else 0.0], dtype=torch.float64, device=dev)
    pc = torch.tensor([1.0 if pn > 0 else 0.0, float(pn)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(pt, op=dist.ReduceOp.MAX)
        dist.all_reduce(pc, op=dist.ReduceOp.SUM)
    parity = {"max_err": float(pt[0].item()), "e2e_max_err": float(pt[1].item()) if e2e is not None else None,
              "n_sample": int(pc[1].item()), "ranks_checked": int(pc[0].item()), "tolerance": PARITY_TOL,
              "oracle": "oracle/_ref (unmodified reference C)" if pkind == "reference" else "oracle port",
              "what": "outputs of the last timed launch on every rank's shard, i-sample x full j-set, "
                      "max over particles of ||d acc||/||acc|| and ||d jerk||/||jerk||",
              "ok": bool(pt.max().item() <= PARITY_TOL)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -----------------------------------------------------
    local_pairs = float(ni) * n
    kern_ms = pair_ms if pair_ms > 0 else ms_per_step
    achieved_tf = FLOPS_PER_PAIR * local_pairs / (kern_ms * 1e-3) * 1e-12
    sm_count = lib.tupan_cuda_sm_count()
    # three denominators: nominal lanes x clock; the best instruction-shape probe
    # (DADD / DMUL / 2-register DFMA run at the nominal rate); the 3-register DFMA chain probe
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    nominal_tf = sm_count * 64 * 2 * 1965e6 * 1e-12
    nominal_sampled_tf = sm_count * 64 * 2 * sm_mhz * 1e6 * 1e-12
    probes = {"dfma_3reg_chain": peak_tf}
    for kind, name in ((1, "dadd"), (2, "dmul"), (3, "dfma_2reg")):
        tops = ctypes.c_double()
        if lib.tupan_cuda_pipe_probe(kind, 100.0, ctypes.byref(tops)) == 0:
            probes[name] = 2.0 * tops.value       # as TFLOP/s-equivalent: 2 flop per FMA lane
    best_probe_tf = max(probes.values())
    algorithmic = int((8 + 6) * ni * 8 + 8 * n * 8)
    chunks = plan[2].value
    workspace = (chunks * 6 * 8 * ni * 2) if chunks > 1 else 0
    roofline = {
        "bound": "fp64_fma", "achieved": achieved_tf, "peak": nominal_tf, "unit": "TFLOP/s",
        "frac": achieved_tf / nominal_tf,
        "frac_of_nominal": achieved_tf / nominal_tf,
        "frac_of_nominal_at_sampled_clock": achieved_tf / nominal_sampled_tf,
        "frac_of_best_probe": achieved_tf / best_probe_tf,
        "frac_of_dfma_chain_probe": achieved_tf / peak_tf,
        "peak_probes_tflops": probes,
        "peak_source": "nominal: %d SMs x 64 FP64 lanes x 2 flop x 1965 MHz = %.2f TFLOP/s (MEASURED_PEAKS.json has "
                       "no FP64 entry); probes measured in this process: DADD / DMUL / 2-register DFMA reach it, a "
                       "chain of 3-register DFMAs needs a third clock per instruction (tools/microbench2.cu)"
                       % (sm_count, nominal_tf),
        "traffic": NCU_TRAFFIC.get((n, world, chunks)),
        "traffic_note": "ncu dram bytes are quoted only for the launch shape they were captured with (n, GPUs, "
                        "j-chunks) = %s; this run: %d chunk(s), accumulator workspace %.2f GB written + read on top "
                        "of %.1f MB algorithmic, %.3f %% of the kernel time at the measured HBM bandwidth"
                        % (sorted(NCU_TRAFFIC), chunks, workspace / 1e9, algorithmic / 1e6,
                           100 * (workspace + algorithmic) / 6.5e12 / (kern_ms * 1e-3)),
        "kernel": KERNEL_NAME, "kernel_ms": kern_ms,
        "kernel_ms_what": "mean over the timed steps of the summed pair-kernel launches of a step (%.1f per step), "
                          "CUDA events on the launching stream" % pair_launches_per_step
        if pair_ms > 0 else "step time (no per-kernel events)",
        "flops_per_pair": FLOPS_PER_PAIR, "pairs_per_launch": local_pairs / max(pair_launches_per_step, 1.0),
        "dp_instr_per_pair": DP_INSTR_PER_PAIR,
        "dp_pipe_frac": DP_INSTR_PER_PAIR * 2 * local_pairs / (kern_ms * 1e-3) * 1e-12 / nominal_tf,
        "algorithmic_bytes": algorithmic,
        "hbm_note": "arithmetic intensity ~2.5e5 flop/B: HBM (%.0f GB/s measured) is not the bound"
                    % json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 0)
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else "HBM is not the bound",
    }

    # ---- CPU baseline on a bounded sample -------------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        clib, kind = cpu_reference_lib()
        cores = os.cpu_count() or 1
        # plain numpy arrays again (the oracle does not care about pinning)
        nis = cpu_sample_size(n, cores, seconds=12.0)
        dt, pairs = cpu_sample(ps, n, clib, cores, nis)
        nis1 = cpu_sample_size(n, 1, seconds=4.0)
        dt1, pairs1 = cpu_sample(ps, n, clib, 1, nis1)
        cpu = {"value": pairs / dt, "unit": "pairs/s", "cores": cores, "kind": kind,
               "sample": "ni=%d evenly spaced i-particles x nj=%d (%.1f s), the unmodified C function on %d "
                         "contiguous i-slices" % (nis, n, dt, cores),
               "single_thread_value": pairs1 / dt1}

    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n),
        "run": {"parallelism": "i-shard x%d, %s" % (
                    world, "j rows read in place over NVLink (peer mappings)"
                    if world > 1 and sk.transport == "p2p" else "j all-gather"),
                "l2": "256 MiB buffer written between steps (inside the timed region)",
                "plan": {"lane_split": plan[0].value, "js_log2": plan[1].value, "jg": plan[2].value}},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "parity": parity,
        "roofline": roofline, "cpu_baseline": cpu, "tflops": FLOPS_PER_PAIR * value * 1e-12,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
