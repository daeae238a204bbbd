"""BASELINE.json configs[2]: Plummer N = 262144, Hermite6 with block time-steps, i-sharded over the
GPUs of one node (one process per GPU under torchrun; replicated state, sharded active set:
tupan_b200/block.py).  Prints one JSON line: block steps, particle steps, pair interactions,
relative energy error, wall time, and a sampled force parity check of the state it ends with
(acc, jerk, snap, crackle of 64 particles against the oracle with the full j-set, 1e-12).

    python tools/run_block.py n=262144 order=6 eta=0.015625 dt_max=0.0078125 t_end=0.0078125
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        tools/run_block.py n=262144 order=6 t_end=0.0078125
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tupan_b200 import backend, ics  # noqa: E402
from tupan_b200.block import BlockHermite  # noqa: E402


def main():
    opts = {"n": 262144, "order": 6, "eta": 1.0 / 64, "dt_max": 2.0 ** -7, "t_end": 2.0 ** -7, "seed": 1,
            "max_steps": 0, "parity": 1}
    for tok in sys.argv[1:]:
        k, v = tok.split("=", 1)
        if k not in opts:
            raise SystemExit("unknown option %r (known: %s)" % (k, ", ".join(sorted(opts))))
        opts[k] = int(v) if k in ("n", "order", "seed", "max_steps", "parity") else float(v)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = backend.require_gpu("float64")
    n = opts["n"]
    ps = ics.make_plummer(n, seed=opts["seed"])
    t0 = time.perf_counter()
    b = BlockHermite(opts["eta"], ps, order=opts["order"], dt_max=opts["dt_max"], device=dev)
    torch.cuda.synchronize()
    t_start = time.perf_counter() - t0
    ke0, pe0 = b.energies()
    l0 = lib.tupan_cuda_launch_count()
    pairs0 = b.pairs
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    hist = {}
    while b.t < opts["t_end"] and (opts["max_steps"] <= 0 or b.block_steps < opts["max_steps"]):
        na = b.step()
        k = int(np.ceil(np.log2(max(na, 1))))
        hist[k] = hist.get(k, 0) + 1
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    launches = lib.tupan_cuda_launch_count() - l0
    ke1, pe1 = b.energies() if b.t >= opts["t_end"] else (float("nan"), float("nan"))
    my_pairs = b.pairs - pairs0
    tot = torch.tensor([my_pairs], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    line = {
        "config": "Plummer N=%d BlockHermite(order=%d) eta=%g dt_max=%g t_end=%g fp64" % (
            n, opts["order"], opts["eta"], opts["dt_max"], opts["t_end"]),
        "n_gpus": world, "block_steps": b.block_steps, "particle_steps": b.particle_steps,
        "mean_active_fraction": b.particle_steps / float(max(b.block_steps, 1) * n),
        "t": b.t, "wall_s": wall, "start_s": t_start, "block_steps_per_s": b.block_steps / wall,
        "particle_steps_per_s": b.particle_steps / wall, "pairs_total": float(tot.item()),
        "pairs_per_s": float(tot.item()) / wall, "pairs_this_rank": my_pairs,
        "gpu_launches_per_block_step": launches / float(max(b.block_steps, 1)),
        "eerr": ((ke1 + pe1) - (ke0 + pe0)) / (-pe1), "ke0": ke0, "pe0": pe0,
        "active_histogram_log2": {str(k): v for k, v in sorted(hist.items())},
    }
    # sampled parity of the derivatives the final state holds: recompute them for a sample with the oracle
    if opts["parity"] and rank == 0:
        import oracle
        from util import rel_err
        S = {k: b.ops.download(b.S[i]) for i, k in enumerate(b.snames)}
        T = b.ops.download(b.T)
        if np.all(T[0] == T[0][0]):                 # synchronous state
            rng = np.random.default_rng(3)
            idx = np.sort(rng.choice(n, 64, replace=False))
            S8 = ("mass", "rx", "ry", "rz", "eps2", "vx", "vy", "vz")
            olib = oracle.load("ref" if oracle.have("ref", "float64") else "oracle", "float64")
            ia = [np.ascontiguousarray(S[a][idx]) for a in S8]
            ja = [np.ascontiguousarray(S[a]) for a in S8]
            ref = [np.zeros(len(idx)) for _ in range(6)]
            oracle.call_threaded(olib, "acc_jerk_kernel", "float64", os.cpu_count() or 1,
                                 *([len(idx)] + ia + [n] + ja + ref))
            # the product path on the same (evolved) state: the rectangular call a block step makes
            st = b._view(b.S, b.snames)
            sel = torch.as_tensor(idx, device=dev)
            ips = b._view(b.ops.take(b.S, sel), b.snames)
            out = b.ops.rows(6, len(idx))
            b.ops.force("acc_jerk_kernel", ips, st, (), [out[k] for k in range(6)])
            got = [b.ops.download(out[k]) for k in range(6)]
            line["parity_acc_jerk"] = rel_err("acc_jerk_kernel", got, ref)
            line["parity_sample"] = len(idx)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
