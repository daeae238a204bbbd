"""Hermite integration with INDIVIDUAL block time-steps on device-resident state
(SURVEY.md 8f, row N2).

tupan's adaptive Hermite (``integrator/hermite.py:343-401``) advances every particle with the
shared minimum block step: one close pair makes all N particles take its step, and every step
costs full N x N force evaluations.  Here each particle carries its own time and its own
power-of-two step (Makino & Aarseth 1992 block scheme, with tupan's ingredients):

* the next block time is ``min_i (time_i + dt_i)``; the particles that reach it are *active*;
* ALL particles are predicted to that time from the derivatives they hold
  (``tupan_cuda_block_predict_dev``);
* forces are evaluated for the active particles only, against everybody's predicted state --
  the same ``acc_jerk`` / ``snap_crackle`` / ``tstep`` kernels, called rectangular (ni = active,
  nj = N), which the launch-plan model turns into lane-split or chunked launches;
* the active particles are corrected with the Hermite corrector of ``hermite.py:93-121``
  (order 4) / ``:160-196`` (order 6) over their own step (``tupan_cuda_block_correct_dev``),
  ``pec`` times as in tupan's ``epec(2, ...)`` with the corrected particles replacing their
  predicted selves in the j-set;
* their new step is the largest power of two not above tupan's pairwise criterion
  (``tstep_kernel``, the ``tstep`` output), at most twice the old one, commensurate with the
  block time, and at most ``dt_max``.

All particles are synchronous at every multiple of ``dt_max``; ``evolve(t_end)`` stops there.

Multi-GPU (BASELINE.json configs[2]: Hermite6, block time-steps, i-sharded over 8 B200; one process
per GPU, ``torch.distributed`` initialised): the state is REPLICATED -- every rank holds all N
particles at their own times and predicts all of them itself (O(N), no communication) -- and the
ACTIVE set of a block step is sharded: every rank derives the same ordered active list from its
replica, takes a contiguous slice of it (so the load is balanced however the active particles are
distributed in index), evaluates forces for its slice against all N predicted particles,
corrects its slice, and the slices are all-gathered (NCCL over NVLink; a few doubles per active
particle: r v a j [s] after every corrector pass, then the new steps) so that every replica
applies the same update.  ``pec + 1`` small all-gathers per block step; no reduction of outputs.
The bookkeeping (minimum, active mask, gather / scatter of the active set) is O(N) array plumbing
done with torch on the device -- per-particle quantities are the rows of 2-D tensors, so the active
subset of a whole state is one ``index_select`` and goes back with one ``index_copy_`` --; the
arithmetic (prediction, forces, corrector, criterion, step quantisation) is kernels of this library.  ``ops`` is the seam the CPU tests use to run the same driver on numpy
arrays with the oracle kernels (oracle/block_ops.py).
"""
import ctypes

import numpy as np

R3, V3, A3, J3, S3 = (("rx", "ry", "rz"), ("vx", "vy", "vz"), ("ax", "ay", "az"), ("jx", "jy", "jz"),
                      ("sx", "sy", "sz"))
S8 = ("mass", "rx", "ry", "rz", "eps2", "vx", "vy", "vz")


class CudaOps(object):
    """Arrays are torch CUDA tensors; arithmetic is the C ABI of include/libtupan_cuda.h.

    Per-particle quantities live as the ROWS of 2-D tensors, so that taking the active subset of
    a whole particle state is one ``index_select`` and putting it back one ``index_copy_``."""

    def __init__(self, device=None):
        import torch
        from . import backend, device as dev
        self.torch, self.dev = torch, dev
        self.device = torch.device(device if device is not None else "cuda")
        self.lib = backend.require_gpu("float64")
        self.backend = backend

    # -- plumbing -------------------------------------------------------------------------
    def rows(self, k, n):
        return self.torch.zeros((k, n), dtype=self.torch.float64, device=self.device)

    def upload(self, dst_row, a):
        dst_row.copy_(self.torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)))

    def download(self, t):
        return t.cpu().numpy()

    def select(self, time, dt):
        """(t_next, ascending indices of the particles that reach it): one launch, one 16-byte read-back."""
        n = time.numel()
        if getattr(self, "_sel_n", None) != n:
            self._sel_n = n
            self._sel_idx = self.torch.empty(n, dtype=self.torch.int64, device=self.device)
            self._sel_out = self.torch.empty(2, dtype=self.torch.float64, device=self.device)
            self._sel_host = self.torch.empty(2, dtype=self.torch.float64).pin_memory()
        self._ok(self.lib.tupan_cuda_block_select_dev(n, ctypes.c_void_p(time.data_ptr()), ctypes.c_void_p(dt.data_ptr()),
                                                      ctypes.c_void_p(self._sel_out.data_ptr()),
                                                      ctypes.c_void_p(self._sel_idx.data_ptr()), self._stream()),
                 "block_select")
        self._sel_host.copy_(self._sel_out, non_blocking=True)
        self.torch.cuda.current_stream().synchronize()
        t_next, na = float(self._sel_host[0]), int(self._sel_host[1])
        return t_next, self._sel_idx[:na]

    def count(self, idx):
        return int(idx.numel())

    def part(self, idx, lo, hi):
        return idx[lo:hi]

    def gather(self, buf, world, group=None):
        """buf [rows, chunk] of every rank -> [rows, world * chunk] (rank-major columns)."""
        if world == 1:
            return buf
        import torch.distributed as dist
        rows, chunk = buf.shape
        out = self.torch.empty(world * rows * chunk, dtype=buf.dtype, device=buf.device)
        dist.all_gather_into_tensor(out, buf.contiguous().view(-1), group=group)
        return out.view(world, rows, chunk).permute(1, 0, 2).reshape(rows, world * chunk)

    def pad(self, block, chunk):
        """[rows, k] -> [rows, chunk] (k <= chunk), zero-filled."""
        rows, k = block.shape
        if k == chunk:
            return block
        out = self.torch.zeros((rows, chunk), dtype=block.dtype, device=block.device)
        out[:, :k] = block
        return out

    def cat(self, blocks):
        return self.torch.cat(blocks, 0)

    def take(self, block, idx):
        return block.index_select(1, idx)

    def put(self, block, idx, src):
        block.index_copy_(1, idx, src)

    # -- arithmetic -----------------------------------------------------------------------
    def _ptrs(self, tensors):
        return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def _ok(self, rc, what):
        if rc != 0:
            self.backend.check(self.lib, what)
            raise self.backend.TupanCudaError("%s failed with code %d" % (what, rc))

    def force(self, kernel, ips, jps, scalars, out_rows):
        self.dev.run(kernel, ips, jps, scalars, dict(zip(self.dev.KERNEL_OUTPUTS[kernel], out_rows)))

    def predict(self, order, state_rows, time, t_next, pred_rows):
        self._ok(self.lib.tupan_cuda_block_predict_dev(order, time.numel(), self._ptrs(state_rows),
                                                       ctypes.c_void_p(time.data_ptr()), float(t_next),
                                                       self._ptrs(pred_rows), self._stream()), "block_predict")

    def correct(self, order, tau, rv0, d0, d1, rv):
        self._ok(self.lib.tupan_cuda_block_correct_dev(order, tau.numel(), ctypes.c_void_p(tau.data_ptr()),
                                                       self._ptrs(rv0), self._ptrs(d0), self._ptrs(d1),
                                                       self._ptrs(rv), self._stream()), "block_correct")

    def quantize(self, ts, tau, t_next, dt_max, dt_new, time_new):
        self._ok(self.lib.tupan_cuda_block_quantize_dev(ts.numel(), ctypes.c_void_p(ts.data_ptr()),
                                                        ctypes.c_void_p(tau.data_ptr()), float(t_next),
                                                        float(dt_max), ctypes.c_void_p(dt_new.data_ptr()),
                                                        ctypes.c_void_p(time_new.data_ptr()), self._stream()),
                 "block_quantize")


def _rows(block):
    return [block[k] for k in range(block.shape[0])]


class BlockHermite(object):
    """``BlockHermite(eta, ps, order=4).evolve(t_end)``; ``ps`` is a particle container with the
    reference's attribute names (mass, eps2, rx.., vx..)."""

    def __init__(self, eta, ps, order=4, dt_max=2.0 ** -3, pec=2, t0=0.0, ops=None, device=None, group=None):
        if order not in (4, 6):
            raise ValueError("order 4 or 6")
        self.eta, self.order, self.dt_max, self.pec = float(eta), int(order), float(dt_max), int(pec)
        self.ops = o = ops or CudaOps(device)
        # one process per GPU: replicated state, the active set of each block step is sharded
        self.group = group
        self.rank, self.world = 0, 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        except ImportError:
            pass
        self.pairs = 0.0                              # pair interactions evaluated by THIS rank
        self.n = n = int(len(ps.mass))
        self.nd = order // 2                                                   # derivative levels: a j (s)
        levels = (R3, V3, A3, J3) + ((S3,) if order >= 6 else ())             # what a particle holds
        self.snames = ("mass", "eps2") + tuple(k for lev in levels for k in lev)
        self.pnames = ("mass", "eps2") + R3 + V3 + (A3 + J3 if order >= 6 else ())
        self.S = o.rows(len(self.snames), n)          # state at each particle's own time
        self.P = o.rows(len(self.pnames), n)          # everybody predicted to the current block time
        self.T = o.rows(2, n)                         # time, dt
        for k in ("mass", "eps2") + R3 + V3:
            o.upload(self.S[self.snames.index(k)], getattr(ps, k))
        for k in ("mass", "eps2"):
            o.upload(self.P[self.pnames.index(k)], getattr(ps, k))
        o.upload(self.T[0], np.full(n, float(t0)))
        self.t = float(t0)
        self.block_steps = 0
        self.particle_steps = 0
        self._start()

    def _view(self, block, names):
        return dict(zip(names, _rows(block)))

    def _derivs(self, ips, jps, d1, scratch):
        """a, j (and s) of the i-set against the j-set into the rows of d1."""
        o = self.ops
        o.force("acc_jerk_kernel", ips, jps, (), _rows(d1)[:6])
        if self.order >= 6:
            ii = dict(ips)
            ii.update(zip(A3 + J3, _rows(d1)[:6]))     # the i side uses the a, j just computed
            o.force("snap_crackle_kernel", ii, jps, (), _rows(d1)[6:9] + _rows(scratch))

    def _slice(self, na):
        """This rank's contiguous part [lo, hi) of an ordered list of na items, and the common chunk."""
        chunk = -(-na // self.world)
        lo = min(self.rank * chunk, na)
        return lo, min(lo + chunk, na), chunk

    def _exchange(self, mine, chunk, na):
        """All ranks' slices of per-active-particle rows, in active-list order: [rows, na]."""
        o = self.ops
        if self.world == 1:
            return mine
        return o.gather(o.pad(mine, chunk), self.world, self.group)[:, :na]

    def _start(self):
        o, n = self.ops, self.n
        st = self._view(self.S, self.snames)
        # every particle is active at the start: the same sharded evaluation as in a block step
        lo, hi, chunk = self._slice(n)
        mine = o.take(self.S, self._arange(lo, hi))
        ips = self._view(mine, self.snames)
        d1, scratch = o.rows(3 * self.nd, hi - lo), o.rows(3, hi - lo)
        self._derivs(ips, st, d1, scratch)
        self.pairs += float(hi - lo) * n * (2 if self.order >= 6 else 1)
        if self.order >= 6:
            # snap needs a, j of ALL particles on the j side: exchange them, then evaluate again
            aj = self._exchange(d1[0:6], chunk, n)
            self.S[8:14] = aj
            ips = self._view(o.take(self.S, self._arange(lo, hi)), self.snames)
            self._derivs(ips, st, d1, scratch)
            self.pairs += float(hi - lo) * n * 2
        ts = o.rows(2, hi - lo)
        o.force("tstep_kernel", ips, st, (self.eta,), _rows(ts))
        self.pairs += float(hi - lo) * n
        # first step: the largest power of two <= criterion and <= dt_max (tau = dt_max/2, t = 0
        # lets block_quantize go up to dt_max)
        tau = o.rows(1, hi - lo)
        o.upload(tau[0], np.full(hi - lo, self.dt_max / 2))
        new = o.rows(2, hi - lo)
        o.quantize(ts[0], tau[0], 0.0, self.dt_max, new[1], new[0])
        allrows = self._exchange(o.cat([d1, new[1:2]]), chunk, n)
        self.S[8:8 + 3 * self.nd] = allrows[:3 * self.nd]
        self.T[1] = allrows[3 * self.nd]

    def _arange(self, lo, hi):
        o = self.ops
        if hasattr(o, "torch"):
            return o.torch.arange(lo, hi, device=o.device)
        return np.arange(lo, hi)

    def step(self):
        """One block step: returns the number of particles advanced (over all ranks)."""
        o, nd = self.ops, self.nd
        S, P, T = self.S, self.P, self.T
        t_next, idx = o.select(T[0], T[1])
        na = o.count(idx)
        o.predict(self.order, _rows(S)[2:], T[0], t_next, _rows(P)[2:])
        lo, hi, chunk = self._slice(na)
        mine = o.part(idx, lo, hi)
        nm = hi - lo
        Sa, Pa, Ta = o.take(S, mine), o.take(P, mine), o.take(T, mine)
        tau = Ta[1]
        d1, scratch = o.rows(3 * nd, nm), o.rows(3, nm)
        ips, jps = self._view(Pa, self.pnames), self._view(P, self.pnames)
        rv0, d0 = _rows(Sa)[2:8], _rows(Sa)[8:8 + 3 * nd]
        for _ in range(self.pec):
            if nm > 0:
                self._derivs(ips, jps, d1, scratch)
                o.correct(self.order, tau, rv0, d0, _rows(d1), _rows(Pa)[2:8])
            self.pairs += float(nm) * self.n * (2 if self.order >= 6 else 1)
            # the corrected particles (of every rank) replace their predicted selves in the j-set
            upd = self._exchange(o.cat([Pa[2:8], d1]), chunk, na)
            o.put(P[2:8], idx, upd[0:6])
            if self.order >= 6:
                o.put(P[8:14], idx, upd[6:12])
        ts, new = o.rows(2, nm), o.rows(2, nm)
        if nm > 0:
            o.force("tstep_kernel", ips, jps, (self.eta,), _rows(ts))
            o.quantize(ts[0], tau, t_next, self.dt_max, new[1], new[0])
        self.pairs += float(nm) * self.n
        new_all = self._exchange(new, chunk, na)
        o.put(S[2:8], idx, upd[0:6])
        o.put(S[8:8 + 3 * nd], idx, upd[6:6 + 3 * nd])
        o.put(T, idx, new_all)
        self.t = t_next
        self.block_steps += 1
        self.particle_steps += na
        return na

    def evolve(self, t_end):
        """Advance to ``t_end`` (a multiple of dt_max: every particle is synchronous there)."""
        if abs(t_end / self.dt_max - round(t_end / self.dt_max)) > 1e-12:
            raise ValueError("t_end must be a multiple of dt_max = %g" % self.dt_max)
        while self.t < t_end:
            self.step()
        return self.block_steps

    def download(self, ps):
        """Write positions, velocities, times and steps back into a host container."""
        o = self.ops
        for k in R3 + V3:
            getattr(ps, k)[...] = o.download(self.S[self.snames.index(k)])
        ps.time[...] = o.download(self.T[0])
        ps.tstep[...] = o.download(self.T[1])
        return ps

    def energies(self):
        """(kinetic, potential) of the synchronous state, phi from the phi kernel."""
        o = self.ops
        st = self._view(self.S, self.snames)
        phi = o.rows(1, self.n)
        o.force("phi_kernel", st, st, (), _rows(phi))
        m = o.download(st["mass"])
        v2 = sum(o.download(st[k]) ** 2 for k in V3)
        return float(0.5 * np.sum(m * v2)), float(0.5 * np.sum(m * o.download(phi[0])))
