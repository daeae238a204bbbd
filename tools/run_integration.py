"""Run one of BASELINE.json's integration configs on the device-resident integrators and print
one JSON line: steps, energies, relative energy error (te - te0)/(-pe) (the reference's
Diagnostic, simulation.py:109-112), wall time, steps/s and pair-interactions/s.

    python tools/run_integration.py n=1024 method=ahermite4 eta=0.015625 t_end=1
    python tools/run_integration.py n=65536 method=sia21s.dkd eta=0.00390625 t_end=0.015625 prec=float32
    torchrun ... tools/run_integration.py n=262144 method=ahermite6 max_steps=4      # i-sharded
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tupan_b200 import backend, ics  # noqa: E402
from tupan_b200.integrator import Integrator  # noqa: E402

# pair kernels evaluated per step (N x N pairs each)
def evals_per_step(method):
    if "hermite" in method:
        o = int(method[-1])
        per = 1 if o <= 4 else 2
        return 3 * per + (1 if method.startswith("a") else 0)
    if method.startswith("sia"):
        from tupan_b200.integrator import SIA_COEFS, operator_sequence
        A, B = SIA_COEFS[method[:5]]
        seq = operator_sequence(B, A)
        kicks = sum(1 for outer, _ in seq if outer == method.endswith("kdk"))
        bridges = sum(1 for outer, _ in seq if outer)
        return kicks * bridges + (1 if method[5] == "a" else 0)
    if "sakura" in method:
        return 2 + (1 if method.startswith("a") else 0)
    return 3 if method == "anreg" else 6


def main():
    # key=value tokens on purpose: torchrun's own parser abbreviation-matches long options that
    # follow the script name (--n -> --nnodes ...)
    opts = {"n": 1024, "method": "ahermite4", "eta": 1.0 / 64, "t_end": 1.0, "prec": "float64", "seed": 1,
            "max_steps": None, "check_every": 16, "graph": None, "plan": None}
    for tok in sys.argv[1:]:
        k, v = tok.split("=", 1)
        if k not in opts:
            raise SystemExit("unknown option %r (known: %s)" % (k, ", ".join(sorted(opts))))
        opts[k] = v if k in ("method", "prec", "plan") else (int(v) if k in ("n", "seed", "max_steps", "check_every", "graph")
                                                     else float(v))
    args = argparse.Namespace(**opts)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = backend.require_gpu(args.prec)
    if args.plan:                                # plan=split,js,jg forces the launch shape of every pair kernel
        lib.tupan_cuda_force_plan(*[int(x) for x in args.plan.split(",")])
    ps = ics.make_plummer(args.n, seed=args.seed, dtype=args.prec)
    it = Integrator(args.eta, 0.0, ps, method=args.method, device=dev,
                    graph=None if args.graph is None else bool(args.graph))
    ke0, pe0 = it.energies()
    it.evolve_step(args.t_end)                   # warm-up step: buffers sized, kernels loaded
    torch.cuda.synchronize()
    l0 = lib.tupan_cuda_launch_count()
    s0 = it.nsteps
    t0 = time.perf_counter()
    steps = it.evolve(args.t_end, check_every=args.check_every, max_steps=args.max_steps)
    it.finalize(args.t_end)
    wall = time.perf_counter() - t0
    launches = lib.tupan_cuda_launch_count() - l0
    ke1, pe1 = it.energies()
    timed = steps - s0
    if rank == 0:
        per = evals_per_step(args.method)
        print(json.dumps({
            "config": "Plummer N=%d %s eta=%g t_end=%g %s%s" % (args.n, args.method, args.eta, args.t_end, args.prec,
                                                              " plan=" + args.plan if args.plan else ""),
            "n_gpus": world, "steps": steps, "t": it.time, "ke0": ke0, "pe0": pe0, "ke1": ke1, "pe1": pe1,
            "eerr": ((ke1 + pe1) - (ke0 + pe0)) / (-pe1), "wall_s": wall, "timed_steps": timed,
            "steps_per_s": timed / wall, "us_per_step": 1e6 * wall / max(timed, 1),
            "pair_kernel_evals_per_step": per,
            "pairs_per_s": per * float(args.n) ** 2 * timed / wall,
            "gpu_launches_per_step": launches / max(timed, 1), "cuda_graph": it._graph is not None}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
