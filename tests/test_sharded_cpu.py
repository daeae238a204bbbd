"""CPU tests of the multi-GPU host logic (tupan_b200/sharded.py) with the gloo backend,
world_size 2 and 3: partition, packed-row all-gather layout (incl. uneven shards), slot
bookkeeping and finalize.  The arithmetic seam is filled by a numpy/oracle engine that lives
here in tests/ -- the product engine is CUDA-only."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


class OracleEngine(object):
    """pack = rows [j][n_in] in libtupan.h order; sweep = oracle call on a row range (final
    outputs of that range, valid as raw accumulators for the sum-type kernels); finalize = sum."""

    def __init__(self):
        import oracle
        from util import KERNELS
        self.oracle = oracle
        self.lib = oracle.load("oracle", "float64")
        self.K = KERNELS
        self.chunk_rows = 37          # several slots per sweep: exercises the slot bookkeeping

    def row_width(self, kernel, scal):
        return len(self.K[kernel][0])

    def n_acc(self, kernel, scal):
        return self.oracle.n_outputs(kernel)

    def pack(self, kernel, jt, scal, packed):
        w = self.row_width(kernel, scal)
        rows = torch.stack(jt, dim=1).reshape(-1)
        packed[:rows.numel()] = rows
        assert rows.numel() == jt[0].numel() * w

    def sweep_slots(self, kernel, ni, rows, scal):
        return max(1, -(-rows // self.chunk_rows))

    def sweep(self, kernel, it, packed, j0, j1, scal, partial, slot0):
        w = self.row_width(kernel, scal)
        na = self.n_acc(kernel, scal)
        ni = it[0].numel()
        rows = packed.numpy()[:j1 * w].reshape(-1, w)
        for s in range(self.sweep_slots(kernel, ni, j1 - j0, scal)):
            a, b = j0 + s * self.chunk_rows, min(j1, j0 + (s + 1) * self.chunk_rows)
            ja = [np.ascontiguousarray(rows[a:b, k]) for k in range(w)]
            outs = [np.zeros(ni) for _ in range(na)]
            args = [ni] + [t.numpy() for t in it] + [b - a] + ja + list(scal) + outs
            self.oracle.call(self.lib, kernel, "float64", *args)
            dst = partial.numpy()[(slot0 + s) * na * ni:(slot0 + s + 1) * na * ni].reshape(na, ni)
            for k in range(na):
                dst[k] = outs[k]

    # the rows of several owners in one call (the peer transport's single launch): here simply the
    # owners one after the other
    def sweep_multi_slots(self, kernel, ni, seg_rows, scal):
        return sum(self.sweep_slots(kernel, ni, r, scal) for r in seg_rows if r > 0)

    def sweep_multi(self, kernel, it, seg_ptrs, seg_rows, scal, partial, slot0):
        slot = slot0
        for ptr, r in zip(seg_ptrs, seg_rows):
            if r > 0:
                self.sweep(kernel, it, ptr, 0, r, scal, partial, slot)
                slot += self.sweep_slots(kernel, it[0].numel(), r, scal)

    def finalize(self, kernel, it, partial, nslots, scal, ot):
        na = self.n_acc(kernel, scal)
        ni = it[0].numel()
        acc = partial.numpy()[:nslots * na * ni].reshape(nslots, na, ni).sum(0)
        for k in range(na):
            ot[k].copy_(torch.from_numpy(acc[k]))


class GlooRows(object):
    """Stand-in for sharded.PeerRows on CPU: `rows[r]` are views of a buffer that `barrier()` fills
    with an all-gather -- the same seam (pack into rows[rank], barrier, read rows[r]) without peer
    mappings."""

    def __init__(self, rows_max, width, group=None):
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.chunk = rows_max * width
        self.buf = torch.zeros(self.world * self.chunk, dtype=torch.float64)
        self.rows = [self.buf[r * self.chunk:(r + 1) * self.chunk] for r in range(self.world)]
        self.group = group
        self.barriers = 0

    def barrier(self):
        self.barriers += 1
        dist.all_gather_into_tensor(self.buf, self.rows[self.rank].clone(), group=self.group)

    def check(self):
        pass


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, kernel, q, mode="nccl"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tupan_b200 import ics, sharded
        from util import KERNELS, as_dict
        data = as_dict(ics.make_plummer(n, seed=3), "float64")
        peer = None
        if mode != "nccl":              # the peer transport's host logic: one launch / one launch per owner
            rows_max = max(b - a for a, b in zip(sharded.shard_bounds(n, world)[:-1], sharded.shard_bounds(n, world)[1:]))
            peer = GlooRows(rows_max, sharded.ShardedKernel.MAX_ROW_WIDTH)
        sk = sharded.ShardedKernel(kernel, n, torch.float64, "cpu", engine=OracleEngine(), peer=peer)
        if mode == "p2p-per-owner":
            sk.MULTI_MAX_PAIRS = 0.0
        assert sk.transport == ("nccl" if mode == "nccl" else "p2p")
        local = {a: torch.from_numpy(data[a][sk.lo:sk.hi].copy()) for a in KERNELS[kernel][0]}
        out = sk.evaluate(local)
        out = sk.evaluate(local, (), out)          # second call reuses the buffers
        res = {k: v.numpy().copy() for k, v in out.items()}
        q.put((rank, sk.lo, sk.hi, sk.segments(), res))
    finally:
        dist.destroy_process_group()


CASES = [(w, n, k, "nccl") for (w, n) in ((2, 200), (3, 101), (2, 3)) for k in ("acc_jerk_kernel", "phi_kernel")]
# the peer transport's host logic (rows stay with their owners; one launch, or one per owner)
CASES += [(3, 101, "acc_jerk_kernel", "p2p-one-launch"), (2, 3, "phi_kernel", "p2p-one-launch"),
          (3, 101, "phi_kernel", "p2p-per-owner"), (2, 3, "acc_jerk_kernel", "p2p-per-owner")]


@pytest.mark.parametrize("world,n,kernel,mode", CASES)
def test_sharded_evaluation_equals_single_call(world, n, kernel, mode):
    import oracle
    from tupan_b200 import ics
    from tupan_b200.device import KERNEL_OUTPUTS
    from util import as_dict, run
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, kernel, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    data = as_dict(ics.make_plummer(n, seed=3), "float64")
    ref = run(oracle.load("oracle", "float64"), kernel, "float64", data, data)
    covered = np.zeros(n, bool)
    for rank, lo, hi, segs, res in got:
        covered[lo:hi] = True
        if mode == "nccl":
            assert sum(1 for s in segs if s[2]) == (1 if hi > lo else 0)
            assert sum(j1 - j0 for (j0, j1, _) in segs) == n   # every j row swept exactly once
        for k, name in enumerate(KERNEL_OUTPUTS[kernel]):
            np.testing.assert_allclose(res[name], ref[k][lo:hi], rtol=1e-12, atol=0)
    assert covered.all()


def test_shard_bounds_and_segments():
    from tupan_b200 import sharded
    assert sharded.shard_bounds(10, 4) == [0, 2, 5, 7, 10]
    assert sharded.shard_bounds(1 << 20, 8)[1] == 131072
    sk = sharded.ShardedKernel.__new__(sharded.ShardedKernel)
    sk.world, sk.rank = 4, 1
    sk.bounds = sharded.shard_bounds(16, 4)
    sk.rows_max = 4
    # equal shards: [0,4) remote | [4,8) local | [8,16) remote merged
    assert sk.segments() == [(0, 4, False), (4, 8, True), (8, 16, False)]
    sk.bounds = sharded.shard_bounds(10, 4)
    sk.rows_max = 3
    # uneven shards are padded to rows_max: gaps break the merge
    assert sk.segments() == [(0, 2, False), (3, 6, True), (6, 8, False), (9, 12, False)]
