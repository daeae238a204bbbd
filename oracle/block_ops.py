"""numpy + oracle-kernel ``ops`` for tupan_b200.block.BlockHermite -- TEST INFRASTRUCTURE ONLY.

The reference has no individual-block-step integrator (its adaptive Hermite shares the minimum
block step, integrator/hermite.py:343-401), so there is no reference run to pin this against:
parity unpinned.  What this file provides is the CPU side of a CPU-vs-GPU comparison of the
same driver -- the pairwise arithmetic is the pinned C restatement of the reference kernels
(oracle/tupan_oracle.c), predictor and corrector are restated here in numpy from
hermite.py:75-121 (order 4) and :127-196 (order 6).
"""
import numpy as np

from . import SIGNATURES, call, load

_IN = {"acc_jerk_kernel": ("mass", "rx", "ry", "rz", "eps2", "vx", "vy", "vz"),
       "tstep_kernel": ("mass", "rx", "ry", "rz", "eps2", "vx", "vy", "vz"),
       "phi_kernel": ("mass", "rx", "ry", "rz", "eps2"),
       "snap_crackle_kernel": ("mass", "rx", "ry", "rz", "eps2", "vx", "vy", "vz",
                               "ax", "ay", "az", "jx", "jy", "jz")}


class OracleOps(object):
    def __init__(self, kind="oracle"):
        self.lib = load(kind, "float64")

    def upload(self, a):
        return np.array(a, dtype=np.float64)

    def download(self, a):
        return a

    def full(self, n, value):
        return np.full(n, float(value))

    def next_time(self, time, dt):
        return float((time + dt).min())

    def active(self, time, dt, t_next):
        return np.nonzero((time + dt) == t_next)[0]

    def count(self, idx):
        return int(len(idx))

    def gather(self, a, idx):
        return np.ascontiguousarray(a[idx])

    def scatter(self, a, idx, values):
        a[idx] = values

    def pow2_floor(self, x):
        m, e = np.frexp(x)
        return np.ldexp(np.ones_like(x), e - 1)

    def minimum(self, a, b):
        return np.minimum(a, b)

    def where(self, c, a, b):
        return np.where(c, a, b)

    def remainder_is_zero(self, t, d):
        return np.remainder(np.full_like(d, float(t)), d) == 0

    def force(self, kernel, ips, jps, scalars=()):
        ins = _IN[kernel]
        ni, nj = len(ips[ins[0]]), len(jps[ins[0]])
        outs = [np.zeros(ni) for _ in range(SIGNATURES[kernel].count("O"))]
        args = ([ni] + [np.ascontiguousarray(ips[k]) for k in ins] + [nj]
                + [np.ascontiguousarray(jps[k]) for k in ins] + list(scalars) + outs)
        call(self.lib, kernel, "float64", *args)
        return outs

    def predict(self, order, state, time, t_next):
        # Taylor series of the levels a particle holds (r v a j [s]), Horner form
        nl = order // 2 + 2
        npred = 4 if order >= 6 else 2
        dt = t_next - time
        pred = []
        for m in range(npred):
            for c in range(3):
                x = state[3 * (nl - 1) + c].copy()
                for k in range(nl - 1, m, -1):
                    x = x * dt / (k - m) + state[3 * (k - 1) + c]
                pred.append(x)
        return pred

    @staticmethod
    def _corr(nd, p0, p1, tau):
        if nd == 2:     # hermite.py:103-121
            return ((p0[2] - p1[2]) * tau / 6 + (p0[1] + p1[1])) * tau / 2 + p0[0]
        # hermite.py:170-196
        return (((p0[3] + p1[3]) * tau / 12 + (p0[2] - p1[2])) * tau / 5 + (p0[1] + p1[1])) * tau / 2 + p0[0]

    def correct(self, order, tau, rv0, d0, d1):
        nd = order // 2
        r1, v1 = [], []
        for c in range(3):
            p0 = [rv0[3 + c]] + [d0[3 * q + c] for q in range(nd)]
            p1 = [rv0[3 + c]] + [d1[3 * q + c] for q in range(nd)]
            v = self._corr(nd, p0, p1, tau)
            q0 = [rv0[c], rv0[3 + c]] + p0[1:nd]
            q1 = [rv0[c], v] + p1[1:nd]
            r1.append(self._corr(nd, q0, q1, tau))
            v1.append(v)
        return r1 + v1
