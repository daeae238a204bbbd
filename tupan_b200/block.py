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
The bookkeeping (minimum, active mask, gather / scatter of the active set, step quantisation) is
O(N) array plumbing done with torch on the device; every O(N_active x N) operation is one of the
kernels of this library.  ``ops`` is the seam the CPU tests use to run the same driver on numpy
arrays with the oracle kernels (oracle/block_ops.py).
"""
import ctypes

import numpy as np

R3, V3, A3, J3, S3 = (("rx", "ry", "rz"), ("vx", "vy", "vz"), ("ax", "ay", "az"), ("jx", "jy", "jz"),
                      ("sx", "sy", "sz"))
S8 = ("mass", "rx", "ry", "rz", "eps2", "vx", "vy", "vz")


class CudaOps(object):
    """Arrays are torch CUDA tensors; arithmetic is the C ABI of include/libtupan_cuda.h."""

    def __init__(self, device=None):
        import torch
        from . import backend, device as dev
        self.torch, self.dev = torch, dev
        self.device = torch.device(device if device is not None else "cuda")
        self.lib = backend.require_gpu("float64")
        self.backend = backend

    # -- plumbing -------------------------------------------------------------------------
    def upload(self, a):
        return self.torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=self.device)

    def download(self, t):
        return t.cpu().numpy()

    def full(self, n, value):
        return self.torch.full((n,), float(value), dtype=self.torch.float64, device=self.device)

    def next_time(self, time, dt):
        return float((time + dt).min().item())

    def active(self, time, dt, t_next):
        return self.torch.nonzero((time + dt) == t_next).flatten()

    def count(self, idx):
        return int(idx.numel())

    def gather(self, a, idx):
        return a.index_select(0, idx)

    def scatter(self, a, idx, values):
        a.index_copy_(0, idx, values)

    def pow2_floor(self, x):
        m, e = self.torch.frexp(x)                       # x = m 2^e, m in [0.5, 1)
        return self.torch.ldexp(self.torch.ones_like(x), e - 1)

    def minimum(self, a, b):
        return self.torch.minimum(a, b if hasattr(b, "shape") else self.torch.full_like(a, float(b)))

    def where(self, c, a, b):
        return self.torch.where(c, a, b)

    def remainder_is_zero(self, t, d):
        return self.torch.remainder(self.torch.full_like(d, float(t)), d) == 0

    # -- arithmetic -----------------------------------------------------------------------
    def _ptrs(self, tensors):
        return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def _ok(self, rc, what):
        if rc != 0:
            self.backend.check(self.lib, what)
            raise self.backend.TupanCudaError("%s failed with code %d" % (what, rc))

    def force(self, kernel, ips, jps, scalars=()):
        out = self.dev.run(kernel, ips, jps, scalars)
        return [out[k] for k in self.dev.KERNEL_OUTPUTS[kernel]]

    def predict(self, order, state, time, t_next):
        n = time.numel()
        npred = 12 if order >= 6 else 6
        pred = [self.torch.empty(n, dtype=self.torch.float64, device=self.device) for _ in range(npred)]
        self._ok(self.lib.tupan_cuda_block_predict_dev(order, n, self._ptrs(state), ctypes.c_void_p(time.data_ptr()),
                                                       float(t_next), self._ptrs(pred), self._stream()),
                 "block_predict")
        return pred

    def correct(self, order, tau, rv0, d0, d1):
        n = tau.numel()
        rv = [self.torch.empty(n, dtype=self.torch.float64, device=self.device) for _ in range(6)]
        self._ok(self.lib.tupan_cuda_block_correct_dev(order, n, ctypes.c_void_p(tau.data_ptr()), self._ptrs(rv0),
                                                       self._ptrs(d0), self._ptrs(d1), self._ptrs(rv),
                                                       self._stream()), "block_correct")
        return rv


class BlockHermite(object):
    """``BlockHermite(eta, ps, order=4).evolve(t_end)``; ``ps`` is a particle container with the
    reference's attribute names (mass, eps2, rx.., vx..)."""

    def __init__(self, eta, ps, order=4, dt_max=2.0 ** -3, pec=2, t0=0.0, ops=None, device=None):
        if order not in (4, 6):
            raise ValueError("order 4 or 6")
        self.eta, self.order, self.dt_max, self.pec = float(eta), int(order), float(dt_max), int(pec)
        self.ops = ops or CudaOps(device)
        self.n = int(len(ps.mass))
        o = self.ops
        self.st = {k: o.upload(getattr(ps, k)) for k in S8}
        self.levels = (R3, V3, A3, J3) + ((S3,) if order >= 6 else ())     # what a particle holds
        self.time = o.full(self.n, t0)
        self.t = float(t0)
        self.block_steps = 0
        self.particle_steps = 0
        self._start()

    # derivative names of this order: a j (s)
    @property
    def dnames(self):
        return tuple(k for lev in self.levels[2:] for k in lev)

    def _derivs(self, ips, jps):
        """a, j (and s) of the i-set against the j-set; both dicts hold mass eps2 r v (+ a j on
        the j side when order 6 -- the i side gets the a, j just computed)."""
        o = self.ops
        aj = o.force("acc_jerk_kernel", ips, jps)
        if self.order < 6:
            return aj
        ii = dict(ips)
        ii.update(zip(A3 + J3, aj))
        sc = o.force("snap_crackle_kernel", ii, jps)
        return aj + sc[:3]

    def _start(self):
        o, st = self.ops, self.st
        jps = dict(st)
        aj = o.force("acc_jerk_kernel", st, st)
        st.update(zip(A3 + J3, aj))
        if self.order >= 6:
            sc = o.force("snap_crackle_kernel", st, st)
            st.update(zip(S3, sc[:3]))
        ts = o.force("tstep_kernel", {k: st[k] for k in S8}, {k: st[k] for k in S8}, (self.eta,))[0]
        self.dt = o.minimum(o.pow2_floor(ts), self.dt_max)
        del jps

    def step(self):
        """One block step: returns the number of particles advanced."""
        o, st = self.ops, self.st
        t_next = o.next_time(self.time, self.dt)
        state = [st[k] for lev in self.levels for k in lev]
        pred = o.predict(self.order, state, self.time, t_next)
        pnames = R3 + V3 + (A3 + J3 if self.order >= 6 else ())
        jps = {"mass": st["mass"], "eps2": st["eps2"]}
        jps.update(zip(pnames, pred))
        idx = o.active(self.time, self.dt, t_next)
        tau = o.gather(self.dt, idx)
        rv0 = [o.gather(st[k], idx) for k in R3 + V3]
        d0 = [o.gather(st[k], idx) for k in self.dnames]
        ips = {k: o.gather(jps[k], idx) for k in S8}
        d1 = None
        for _ in range(self.pec):
            d1 = self._derivs(ips, jps)
            rv1 = o.correct(self.order, tau, rv0, d0, d1)
            # the corrected particles replace their predicted selves, as i and as j
            for k, v in zip(R3 + V3, rv1):
                o.scatter(jps[k], idx, v)
                ips[k] = v
            if self.order >= 6:
                for k, v in zip(A3 + J3, d1[:6]):
                    o.scatter(jps[k], idx, v)
        ts = o.force("tstep_kernel", ips, {k: jps[k] for k in S8}, (self.eta,))[0]
        # largest power of two <= ts, at most 2 tau (and only if the block time allows), <= dt_max
        cand = o.minimum(o.pow2_floor(ts), self.dt_max)
        twice = tau * 2.0
        up = (cand >= twice) & o.remainder_is_zero(t_next, twice)
        dt_new = o.where(up, twice, o.minimum(cand, tau))
        for k, v in zip(R3 + V3, rv1):
            o.scatter(st[k], idx, v)
        for k, v in zip(self.dnames, d1):
            o.scatter(st[k], idx, v)
        o.scatter(self.time, idx, o.full(o.count(idx), t_next))
        o.scatter(self.dt, idx, dt_new)
        self.t = t_next
        self.block_steps += 1
        self.particle_steps += o.count(idx)
        return o.count(idx)

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
            getattr(ps, k)[...] = o.download(self.st[k])
        ps.time[...] = o.download(self.time)
        ps.tstep[...] = o.download(self.dt)
        return ps

    def energies(self):
        """(kinetic, potential) of the synchronous state, phi from the phi kernel."""
        o, st = self.ops, self.st
        five = {k: st[k] for k in ("mass", "rx", "ry", "rz", "eps2")}
        phi = o.download(o.force("phi_kernel", five, five)[0])
        m = o.download(st["mass"])
        v2 = sum(o.download(st[k]) ** 2 for k in V3)
        return float(0.5 * np.sum(m * v2)), float(0.5 * np.sum(m * phi))
