"""i-sharded evaluation of a pairwise kernel over the GPUs of one node (one process per GPU).

Every tupan kernel is ``out[i] = finish(reduce_j pair(i, j))`` with independent i, so the
i-set is partitioned into contiguous ranges, one per rank; outputs stay with their owner
and no output reduction exists.  The only exchange is the j-state: each rank packs its own
shard into rows (``tupan_cuda_pack_dev``), the rows are all-gathered ONCE per evaluation
(NCCL over NVLink/NVSwitch) on a communication stream while the compute stream already
sweeps the local rows; the remote rows are swept when they have landed, and one finalize
combines the raw accumulator slots and applies the kernel's epilogue.  (The reference has no
multi-device path at all; this is new work)

``transport="p2p"`` removes the copy altogether: every rank packs into a buffer its peers have
mapped through CUDA IPC (:class:`PeerRows`), the ranks meet at a device-side barrier, and each
sweep is pointed at the owner's buffer -- the pair kernel's TMA bulk copies pull the tiles over
NVLink/NVSwitch while the previous tiles are being computed (Part 2b of
include/libtupan_cuda.h).  No gathered buffer, no NCCL on the data path.

The collective and the arithmetic are reached through two small seams so the partition /
gather / slot bookkeeping can be tested on CPU with the gloo backend:

* ``engine``  -- pack / sweep / finalize on tensors.  The product engine is
  :class:`CudaEngine` (the C ABI of include/libtupan_cuda.h, no fallback).
* ``group``   -- a torch.distributed process group (nccl on GPUs, gloo in the CPU tests).
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import backend
from .device import KERNEL_INPUTS, KERNEL_OUTPUTS, scal_array


def shard_bounds(n, world):
    """Contiguous i-ranges: rank r owns [b[r], b[r+1])."""
    return [(n * r) // world for r in range(world + 1)]


class CudaEngine(object):
    """Building blocks of Part 2 of include/libtupan_cuda.h on CUDA tensors."""

    def __init__(self, prec):
        self.prec = prec
        self.lib = backend.require_gpu(prec)

    @staticmethod
    def _ptrs(tensors):
        return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])

    @staticmethod
    def _stream():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _ok(self, rc, what):
        if rc != 0:
            backend.check(self.lib, what)
            raise backend.TupanCudaError("%s failed with code %d" % (what, rc))

    def row_width(self, kernel, scal):
        return self.lib.tupan_cuda_row_width(backend.KERNEL_IDS[kernel], scal_array(scal))

    def n_acc(self, kernel, scal):
        return self.lib.tupan_cuda_n_acc(backend.KERNEL_IDS[kernel], scal_array(scal))

    @staticmethod
    def _addr(buf):
        """A tensor or a raw device address (rows that live in a peer's memory)."""
        return ctypes.c_void_p(buf if isinstance(buf, int) else buf.data_ptr())

    def pack(self, kernel, jt, scal, packed):
        self._ok(self.lib.tupan_cuda_pack_dev(backend.KERNEL_IDS[kernel], jt[0].numel(), self._ptrs(jt),
                                              scal_array(scal), self._addr(packed),
                                              self._stream()), "pack")

    def sweep_slots(self, kernel, ni, rows, scal):
        return self.lib.tupan_cuda_sweep_slots(backend.KERNEL_IDS[kernel], ni, rows, scal_array(scal))

    def sweep(self, kernel, it, packed, j0, j1, scal, partial, slot0):
        self._ok(self.lib.tupan_cuda_sweep_dev(backend.KERNEL_IDS[kernel], it[0].numel(), self._ptrs(it),
                                               self._addr(packed), j0, j1, scal_array(scal),
                                               ctypes.c_void_p(partial.data_ptr()), slot0, self._stream()),
                 "sweep")

    def sweep_multi_slots(self, kernel, ni, seg_rows, scal):
        rows = (ctypes.c_longlong * len(seg_rows))(*seg_rows)
        return self.lib.tupan_cuda_sweep_multi_slots(backend.KERNEL_IDS[kernel], ni, len(seg_rows), rows,
                                                     scal_array(scal))

    def sweep_multi(self, kernel, it, seg_ptrs, seg_rows, scal, partial, slot0):
        """One launch over the packed rows of several owners (raw device addresses)."""
        ptrs = (ctypes.c_void_p * len(seg_ptrs))(*[p if isinstance(p, int) else p.data_ptr() for p in seg_ptrs])
        rows = (ctypes.c_longlong * len(seg_rows))(*seg_rows)
        self._ok(self.lib.tupan_cuda_sweep_multi_dev(backend.KERNEL_IDS[kernel], it[0].numel(), self._ptrs(it),
                                                     len(seg_ptrs), ptrs, rows, scal_array(scal),
                                                     ctypes.c_void_p(partial.data_ptr()), slot0, self._stream()),
                 "sweep_multi")

    def finalize(self, kernel, it, partial, nslots, scal, ot):
        self._ok(self.lib.tupan_cuda_finalize_dev(backend.KERNEL_IDS[kernel], it[0].numel(), self._ptrs(it),
                                                  ctypes.c_void_p(partial.data_ptr()), nslots,
                                                  scal_array(scal), self._ptrs(ot), self._stream()),
                 "finalize")


class PeerRows(object):
    """This rank's packed-row buffer and flag block, mapped by every rank of the group, and the
    peers' buffers mapped here (CUDA IPC; one process per GPU on one node)."""

    FLAG_BYTES = 256
    TIMEOUT_S = float(os.environ.get("TUPAN_B200_PEER_TIMEOUT", "20"))   # seconds a barrier waits for a peer

    def __init__(self, lib, nbytes, group=None):
        self.lib = lib
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._own = []
        self._mapped = []

        def alloc(n):
            p, h = ctypes.c_void_p(), ctypes.create_string_buffer(64)
            rc = lib.tupan_cuda_peer_alloc(int(n), ctypes.byref(p), h)
            if rc != 0:
                backend.check(lib, "peer_alloc")
                raise backend.TupanCudaError("peer_alloc failed with code %d" % rc)
            self._own.append(p.value)
            return p.value, h.raw

        rows, hrows = alloc(nbytes)
        flags, hflags = alloc(self.FLAG_BYTES)
        handles = [None] * self.world
        dist.all_gather_object(handles, (hrows, hflags), group=group)
        self.rows = [None] * self.world
        self.flags = [None] * self.world
        for r, (hr, hf) in enumerate(handles):
            if r == self.rank:
                self.rows[r], self.flags[r] = rows, flags
                continue
            for h, dst in ((hr, self.rows), (hf, self.flags)):
                p = ctypes.c_void_p()
                rc = lib.tupan_cuda_peer_open(ctypes.create_string_buffer(h, 64), ctypes.byref(p))
                if rc != 0:
                    backend.check(lib, "peer_open")
                    raise backend.TupanCudaError("peer_open failed with code %d" % rc)
                dst[r] = p.value
                self._mapped.append(p.value)
        self._flag_array = (ctypes.c_void_p * self.world)(*self.flags)
        dist.barrier(group=group)        # nobody signals a flag block that is not mapped everywhere yet

    def barrier(self):
        """All ranks meet on their current streams; asynchronous for the host."""
        rc = self.lib.tupan_cuda_peer_barrier_dev(self._flag_array, self.rank, self.world, self.TIMEOUT_S,
                                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            backend.check(self.lib, "peer_barrier")
            raise backend.TupanCudaError("peer_barrier failed with code %d" % rc)

    def check(self):
        """Synchronise and fail loudly if a barrier gave up waiting for a peer."""
        hits = self.lib.tupan_cuda_peer_timeouts()
        if hits != 0:
            raise backend.TupanCudaError("peer barrier: %d wait(s) timed out (or CUDA error)" % hits)

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)   # no peer may still be reading
        for p in self._mapped:
            self.lib.tupan_cuda_peer_close(ctypes.c_void_p(p))
        dist.barrier(group=self.group)
        for p in self._own:
            self.lib.tupan_cuda_peer_free(ctypes.c_void_p(p))
        self._mapped, self._own = [], []


class ShardedKernel(object):
    """One pairwise kernel, i-sharded over the ranks of ``group``.

    ``evaluate(local)`` takes the rank's own shard (dict of 1-D tensors, the attribute names
    of the reference's particle arrays) and returns the kernel outputs for that shard.
    """

    OVERLAP_MIN_PAIRS = 2.0e9      # ~4 ms of acc_jerk fp64 on a B200
    MAX_ROW_WIDTH = 16             # reals per packed row, widest kernel (snap_crackle: 14, padded)

    # peer transport: below this many pairs per rank the owners' rows are swept in ONE launch (the
    # multi-owner kernel; launch- and tail-bound regime), above it with one launch per owner (the
    # single-buffer kernel is 3-4 % faster per pair: 1000 vs 958 Gpair/s at N = 2^20 on 2 GPUs)
    MULTI_MAX_PAIRS = 2.0e10
    # "auto" transport: p2p below this many pairs per rank.  0 = never: with the round-2 launch
    # plan the all-gather path is at least as fast as reading the rows in place at EVERY size on 2
    # and on 8 GPUs (profiles/r02_sweep_accjerk_fp64_n{2,8}_transports.txt: N = 16384 on 8 GPUs
    # 2130 Gpair/s over NCCL, 2083 with the graph-replayed peer sweep; N = 65536 3751 vs 3641), so
    # the peer transport stays an option (TUPAN_B200_TRANSPORT=p2p) and is not chosen on its own.
    AUTO_P2P_PAIRS = 0.0

    def __init__(self, kernel, n_total, dtype=torch.float64, device="cuda", group=None, engine=None,
                 overlap=True, transport=None, peer=None):
        self.kernel = kernel
        self.n = int(n_total)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bounds = shard_bounds(self.n, self.world)
        self.lo, self.hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.rows_max = max(self.bounds[r + 1] - self.bounds[r] for r in range(self.world))
        self.dtype = dtype
        self.device = torch.device(device)
        self.engine = engine or CudaEngine("float64" if dtype == torch.float64 else "float32")
        self.on_cuda = self.device.type == "cuda"
        self.overlap = overlap and self.on_cuda and self.world > 1
        self.comm_stream = torch.cuda.Stream(device=self.device) if self.overlap else None
        self._packed = None
        self._partial = None
        self._width = None
        # "nccl": all-gather of the packed rows; "p2p": rows stay where they were packed and are
        # read through peer mappings (GPUs of one node only); "auto" (default): p2p with ONE
        # multi-owner launch while the problem is latency-bound (fewer than AUTO_P2P_PAIRS pairs per
        # rank: the all-gather's latency and the extra launches are what such sizes pay for), the
        # all-gather above.  Every rank takes the same decision (it depends on n and the world size).
        self.transport = transport or os.environ.get("TUPAN_B200_TRANSPORT", "auto")
        if self.transport not in ("nccl", "p2p", "auto"):
            raise ValueError("transport must be 'nccl', 'p2p' or 'auto'")
        if self.transport == "auto":
            small = float(self.rows_max) * self.n < self.AUTO_P2P_PAIRS
            self.transport = "p2p" if (small and self.world > 1 and self.on_cuda and peer is None
                                       and os.environ.get("TUPAN_B200_NO_P2P") is None) else "nccl"
        if self.world == 1 or not self.on_cuda:
            self.transport = "nccl"
        self._graph = None
        self._graph_key = None
        self.peer = peer               # injected by the CPU tests (rows gathered with gloo behind the same seam)
        if peer is not None:
            self.transport = "p2p"

    # rows of rank r live at [r * rows_max, r * rows_max + count_r) of the gathered buffer
    def segments(self, split_local=True):
        """Contiguous row ranges of the gathered buffer: [(j0, j1, is_local)], adjacent full
        shards merged so that equal shards give at most three sweeps (one when the local rows
        are not swept separately)."""
        segs = []
        for r in range(self.world):
            cnt = self.bounds[r + 1] - self.bounds[r]
            if cnt == 0:
                continue
            j0 = r * self.rows_max
            local = split_local and r == self.rank
            if segs and not local and not segs[-1][2] and segs[-1][1] == j0:
                segs[-1] = (segs[-1][0], j0 + cnt, False)
            else:
                segs.append((j0, j0 + cnt, local))
        return segs

    def _buffers(self, scal):
        width = self.engine.row_width(self.kernel, scal)
        if self._packed is None or self._width != width:
            self._width = width
            self._packed = torch.zeros(self.world * self.rows_max * width, dtype=self.dtype, device=self.device)
        return self._packed, width

    _peer_cache = {}

    def _peer_rows(self):
        """One mapped buffer per (group, element size, rows) shared by every kernel of the process."""
        if self.peer is None:
            esize = torch.empty(0, dtype=self.dtype).element_size()
            key = (id(self.group), esize, self.rows_max)
            if key not in ShardedKernel._peer_cache:
                nbytes = self.rows_max * self.MAX_ROW_WIDTH * esize
                ShardedKernel._peer_cache[key] = PeerRows(self.engine.lib, nbytes, self.group)
            self.peer = ShardedKernel._peer_cache[key]
        return self.peer

    def evaluate_p2p(self, local, scalars=(), out=None):
        """The same evaluation with the rows left in their owners' memory."""
        ins = KERNEL_INPUTS[self.kernel]
        it = [local[a] for a in ins]
        ni = self.hi - self.lo
        if it[0].numel() != ni:
            raise ValueError("rank %d owns %d particles, got %d" % (self.rank, ni, it[0].numel()))
        if out is None:
            out = {a: torch.empty(ni, dtype=self.dtype, device=self.device) for a in KERNEL_OUTPUTS[self.kernel]}
        ot = [out[a] for a in KERNEL_OUTPUTS[self.kernel]]
        eng = self.engine
        peer = self._peer_rows()
        if eng.row_width(self.kernel, scalars) > self.MAX_ROW_WIDTH:
            raise ValueError("packed row wider than the peer buffer")
        if ni > 0:
            eng.pack(self.kernel, it, scalars, peer.rows[self.rank])
        peer.barrier()                               # every rank's rows are in place
        # own rows first, then the peers round-robin so that no two ranks start on the same GPU; the
        # kernel's TMA ring pulls remote tiles over NVLink while the previous tiles are computed
        order = [(self.rank + k) % self.world for k in range(self.world)]
        seg_rows = [self.bounds[r + 1] - self.bounds[r] for r in order]
        seg_ptrs = [peer.rows[r] for r in order]
        one_launch = float(ni) * self.n < self.MULTI_MAX_PAIRS
        if ni <= 0:
            nslots = []
        elif one_launch:
            nslots = [eng.sweep_multi_slots(self.kernel, ni, seg_rows, scalars)]
        else:
            nslots = [eng.sweep_slots(self.kernel, ni, cnt, scalars) if cnt > 0 else 0 for cnt in seg_rows]
        na = eng.n_acc(self.kernel, scalars)
        need = max(sum(nslots), 1) * na * max(ni, 1)
        if self._partial is None or self._partial.numel() < need:
            self._partial = torch.empty(need, dtype=self.dtype, device=self.device)
        slot = 0
        if ni > 0 and one_launch:
            eng.sweep_multi(self.kernel, it, seg_ptrs, seg_rows, scalars, self._partial, 0)
            slot = nslots[0]
        elif ni > 0:
            for ptr, cnt, ns in zip(seg_ptrs, seg_rows, nslots):
                if cnt > 0:
                    eng.sweep(self.kernel, it, ptr, 0, cnt, scalars, self._partial, slot)
                slot += ns
        peer.barrier()                               # everybody is done reading: rows may be repacked
        if ni > 0:
            eng.finalize(self.kernel, it, self._partial, slot, scalars, ot)
        # A barrier that gave up waiting leaves only a counter behind (the sweep went on over stale
        # rows).  Checking it costs a device synchronisation, so it is done every CHECK_EVERY
        # evaluations here and at every synchronisation point of the integrators
        # (integrator._Lib.check_async); `check()` forces it.
        self._since_check = getattr(self, "_since_check", 0) + 1
        if self._since_check >= self.CHECK_EVERY:
            self.check()
        return out

    CHECK_EVERY = 64

    def check(self):
        """Synchronise and raise if a peer barrier of this process timed out."""
        self._since_check = 0
        if self.peer is not None:
            self.peer.check()

    def evaluate_graphed(self, local, scalars, out):
        """The p2p evaluation replayed from a CUDA graph (pack, device-side barrier, one multi-owner
        sweep, barrier, finalize: five launches whose latency is what a small problem costs).  The
        tensors of `local` and `out` must be the same objects from call to call; the graph is
        re-captured when they change.  Only for transport 'p2p' (no NCCL call inside)."""
        if self.transport != "p2p":
            return self.evaluate(local, scalars, out)
        key = (tuple(t.data_ptr() for t in local.values()), tuple(t.data_ptr() for t in out.values()), tuple(scalars))
        if self._graph is None or self._graph_key != key:
            import gc
            self.evaluate_p2p(local, scalars, out)           # eager once: buffers sized, peers mapped
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            gc.collect()
            lib = self.engine.lib
            before = lib.tupan_cuda_launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.evaluate_p2p(local, scalars, out)
            self._graph_launches = lib.tupan_cuda_launch_count() - before
            lib.tupan_cuda_count_launches(-self._graph_launches)
            self._graph, self._graph_key = g, key
            dist.barrier(group=self.group)                   # every rank has its graph before anyone replays
        self._graph.replay()
        self.engine.lib.tupan_cuda_count_launches(self._graph_launches)
        return out

    def evaluate(self, local, scalars=(), out=None):
        if self.transport == "p2p":
            return self.evaluate_p2p(local, scalars, out)
        ins = KERNEL_INPUTS[self.kernel]
        it = [local[a] for a in ins]
        ni = self.hi - self.lo
        if it[0].numel() != ni:
            raise ValueError("rank %d owns %d particles, got %d" % (self.rank, ni, it[0].numel()))
        if out is None:
            out = {a: torch.empty(ni, dtype=self.dtype, device=self.device) for a in KERNEL_OUTPUTS[self.kernel]}
        ot = [out[a] for a in KERNEL_OUTPUTS[self.kernel]]
        eng = self.engine
        packed, width = self._buffers(scalars)
        chunk = self.rows_max * width
        mine = packed[self.rank * chunk:(self.rank + 1) * chunk]
        eng.pack(self.kernel, it, scalars, mine)

        # Sweeping the local rows while the remote ones are in flight hides the all-gather (tens
        # of microseconds) at the price of one more launch and a smaller, less evenly filled
        # grid per sweep: worth it only when the local sweep is long.
        overlap = self.overlap and float(ni) * (self.hi - self.lo) >= self.OVERLAP_MIN_PAIRS
        work = None
        if self.world > 1:
            if not self.on_cuda:
                # gloo (CPU tests): no in-place aliasing of input and output
                dist.all_gather_into_tensor(packed, mine.clone(), group=self.group)
            elif overlap:
                self.comm_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self.comm_stream):
                    work = dist.all_gather_into_tensor(packed, mine, group=self.group, async_op=True)
            else:
                dist.all_gather_into_tensor(packed, mine, group=self.group)

        segs = self.segments(split_local=overlap or not self.on_cuda)
        segs.sort(key=lambda s: not s[2])            # local rows first: they need no communication
        nslots = [eng.sweep_slots(self.kernel, ni, j1 - j0, scalars) for (j0, j1, _) in segs]
        na = eng.n_acc(self.kernel, scalars)
        need = sum(nslots) * na * max(ni, 1)
        if self._partial is None or self._partial.numel() < need:
            self._partial = torch.empty(need, dtype=self.dtype, device=self.device)
        slot = 0
        for (j0, j1, is_local), ns in zip(segs, nslots):
            if not is_local and work is not None:
                work.wait()                          # compute stream waits for the gathered rows
                work = None
            if ni > 0:
                eng.sweep(self.kernel, it, packed, j0, j1, scalars, self._partial, slot)
            slot += ns
        if work is not None:
            work.wait()
        if ni > 0:
            eng.finalize(self.kernel, it, self._partial, slot, scalars, ot)
        return out
