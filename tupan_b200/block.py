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

    def next_time(self, time, dt):
        return float((time + dt).min().item())

    def active(self, time, dt, t_next):
        return self.torch.nonzero((time + dt) == t_next).flatten()

    def count(self, idx):
        return int(idx.numel())

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

    def __init__(self, eta, ps, order=4, dt_max=2.0 ** -3, pec=2, t0=0.0, ops=None, device=None):
        if order not in (4, 6):
            raise ValueError("order 4 or 6")
        self.eta, self.order, self.dt_max, self.pec = float(eta), int(order), float(dt_max), int(pec)
        self.ops = o = ops or CudaOps(device)
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

    def _start(self):
        o, n = self.ops, self.n
        st = self._view(self.S, self.snames)
        d1 = self.S[8:8 + 3 * self.nd]
        scratch = o.rows(3, n)
        self._derivs(st, st, d1, scratch)
        ts = o.rows(2, n)
        o.force("tstep_kernel", st, st, (self.eta,), _rows(ts))
        # first step: the largest power of two <= criterion and <= dt_max (tau = dt_max/2, t = 0
        # lets block_quantize go up to dt_max)
        tau = o.rows(1, n)
        o.upload(tau[0], np.full(n, self.dt_max / 2))
        new = o.rows(2, n)
        o.quantize(ts[0], tau[0], 0.0, self.dt_max, new[1], new[0])
        o.upload(self.T[1], o.download(new[1]))

    def step(self):
        """One block step: returns the number of particles advanced."""
        o, nd = self.ops, self.nd
        S, P, T = self.S, self.P, self.T
        t_next = o.next_time(T[0], T[1])
        o.predict(self.order, _rows(S)[2:], T[0], t_next, _rows(P)[2:])
        idx = o.active(T[0], T[1], t_next)
        na = o.count(idx)
        Sa, Pa, Ta = o.take(S, idx), o.take(P, idx), o.take(T, idx)
        tau = Ta[1]
        d1, scratch = o.rows(3 * nd, na), o.rows(3, na)
        ips, jps = self._view(Pa, self.pnames), self._view(P, self.pnames)
        rv0, d0 = _rows(Sa)[2:8], _rows(Sa)[8:8 + 3 * nd]
        for _ in range(self.pec):
            self._derivs(ips, jps, d1, scratch)
            o.correct(self.order, tau, rv0, d0, _rows(d1), _rows(Pa)[2:8])
            # the corrected particles replace their predicted selves in the j-set
            o.put(P[2:8], idx, Pa[2:8])
            if self.order >= 6:
                o.put(P[8:14], idx, d1[0:6])
        ts, new = o.rows(2, na), o.rows(2, na)
        o.force("tstep_kernel", ips, jps, (self.eta,), _rows(ts))
        o.quantize(ts[0], tau, t_next, self.dt_max, new[1], new[0])
        o.put(S[2:8], idx, Pa[2:8])
        o.put(S[8:8 + 3 * nd], idx, d1)
        o.put(T, idx, new)
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
