"""i-sharded evaluation of a pairwise kernel over the GPUs of one node (one process per GPU).

Every tupan kernel is ``out[i] = finish(reduce_j pair(i, j))`` with independent i, so the
i-set is partitioned into contiguous ranges, one per rank; outputs stay with their owner
and no output reduction exists.  The only exchange is the j-state: each rank packs its own
shard into rows (``tupan_cuda_pack_dev``), the rows are all-gathered ONCE per evaluation
(NCCL over NVLink/NVSwitch) on a communication stream while the compute stream already
sweeps the local rows; the remote rows are swept when they have landed, and one finalize
combines the raw accumulator slots and applies the kernel's epilogue.  (The reference has no
multi-device path at all; this is new work)

The collective and the arithmetic are reached through two small seams so the partition /
gather / slot bookkeeping can be tested on CPU with the gloo backend:

* ``engine``  -- pack / sweep / finalize on tensors.  The product engine is
  :class:`CudaEngine` (the C ABI of include/libtupan_cuda.h, no fallback).
* ``group``   -- a torch.distributed process group (nccl on GPUs, gloo in the CPU tests).
"""
import ctypes

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

    def pack(self, kernel, jt, scal, packed):
        self._ok(self.lib.tupan_cuda_pack_dev(backend.KERNEL_IDS[kernel], jt[0].numel(), self._ptrs(jt),
                                              scal_array(scal), ctypes.c_void_p(packed.data_ptr()),
                                              self._stream()), "pack")

    def sweep_slots(self, kernel, ni, rows, scal):
        return self.lib.tupan_cuda_sweep_slots(backend.KERNEL_IDS[kernel], ni, rows, scal_array(scal))

    def sweep(self, kernel, it, packed, j0, j1, scal, partial, slot0):
        self._ok(self.lib.tupan_cuda_sweep_dev(backend.KERNEL_IDS[kernel], it[0].numel(), self._ptrs(it),
                                               ctypes.c_void_p(packed.data_ptr()), j0, j1, scal_array(scal),
                                               ctypes.c_void_p(partial.data_ptr()), slot0, self._stream()),
                 "sweep")

    def finalize(self, kernel, it, partial, nslots, scal, ot):
        self._ok(self.lib.tupan_cuda_finalize_dev(backend.KERNEL_IDS[kernel], it[0].numel(), self._ptrs(it),
                                                  ctypes.c_void_p(partial.data_ptr()), nslots,
                                                  scal_array(scal), self._ptrs(ot), self._stream()),
                 "finalize")


class ShardedKernel(object):
    """One pairwise kernel, i-sharded over the ranks of ``group``.

    ``evaluate(local)`` takes the rank's own shard (dict of 1-D tensors, the attribute names
    of the reference's particle arrays) and returns the kernel outputs for that shard.
    """

    OVERLAP_MIN_PAIRS = 2.0e9      # ~4 ms of acc_jerk fp64 on a B200

    def __init__(self, kernel, n_total, dtype=torch.float64, device="cuda", group=None, engine=None,
                 overlap=True):
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

    def evaluate(self, local, scalars=(), out=None):
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
