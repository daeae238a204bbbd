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
    """Same interface as tupan_b200.block.CudaOps on 2-D numpy arrays (one row per quantity)."""

    def __init__(self, kind="oracle"):
        self.lib = load(kind, "float64")

    def rows(self, k, n):
        return np.zeros((k, n))

    def upload(self, dst_row, a):
        dst_row[...] = a

    def download(self, a):
        return np.array(a)

    def select(self, time, dt):
        t_next = float((time + dt).min())
        return t_next, np.nonzero((time + dt) == t_next)[0]

    def count(self, idx):
        return int(len(idx))

    def part(self, idx, lo, hi):
        return idx[lo:hi]

    def gather(self, buf, world, group=None):
        """[rows, chunk] of every rank -> [rows, world * chunk]; gloo in the CPU tests."""
        if world == 1:
            return buf
        import torch
        import torch.distributed as dist
        rows, chunk = buf.shape
        out = torch.empty(world * rows * chunk, dtype=torch.float64)
        dist.all_gather_into_tensor(out, torch.from_numpy(np.ascontiguousarray(buf)).reshape(-1), group=group)
        return out.view(world, rows, chunk).permute(1, 0, 2).reshape(rows, world * chunk).numpy().copy()

    def pad(self, block, chunk):
        rows, k = block.shape
        if k == chunk:
            return block
        out = np.zeros((rows, chunk))
        out[:, :k] = block
        return out

    def cat(self, blocks):
        return np.concatenate(blocks, 0)

    def take(self, block, idx):
        return np.ascontiguousarray(block[:, idx])

    def put(self, block, idx, src):
        block[:, idx] = src

    def force(self, kernel, ips, jps, scalars, out_rows):
        ins = _IN[kernel]
        ni, nj = len(ips[ins[0]]), len(jps[ins[0]])
        assert len(out_rows) == SIGNATURES[kernel].count("O")
        args = ([ni] + [np.ascontiguousarray(ips[k]) for k in ins] + [nj]
                + [np.ascontiguousarray(jps[k]) for k in ins] + list(scalars) + list(out_rows))
        call(self.lib, kernel, "float64", *args)

    def predict(self, order, state_rows, time, t_next, pred_rows):
        # Taylor series of the levels a particle holds (r v a j [s]), Horner form
        nl = order // 2 + 2
        npred = 4 if order >= 6 else 2
        dt = t_next - time
        for m in range(npred):
            for c in range(3):
                x = state_rows[3 * (nl - 1) + c].copy()
                for k in range(nl - 1, m, -1):
                    x = x * dt / (k - m) + state_rows[3 * (k - 1) + c]
                pred_rows[3 * m + c][...] = x

    @staticmethod
    def _corr(nd, p0, p1, tau):
        if nd == 2:     # hermite.py:103-121
            return ((p0[2] - p1[2]) * tau / 6 + (p0[1] + p1[1])) * tau / 2 + p0[0]
        # hermite.py:170-196
        return (((p0[3] + p1[3]) * tau / 12 + (p0[2] - p1[2])) * tau / 5 + (p0[1] + p1[1])) * tau / 2 + p0[0]

    def correct(self, order, tau, rv0, d0, d1, rv):
        nd = order // 2
        for c in range(3):
            p0 = [rv0[3 + c]] + [d0[3 * q + c] for q in range(nd)]
            p1 = [rv0[3 + c]] + [d1[3 * q + c] for q in range(nd)]
            v = self._corr(nd, p0, p1, tau)
            q0 = [rv0[c], rv0[3 + c]] + p0[1:nd]
            q1 = [rv0[c], v] + p1[1:nd]
            rv[c][...] = self._corr(nd, q0, q1, tau)
            rv[3 + c][...] = v

    def quantize(self, ts, tau, t_next, dt_max, dt_new, time_new):
        m, e = np.frexp(ts)
        cand = np.minimum(np.ldexp(np.ones_like(ts), e - 1), dt_max)
        twice = 2.0 * tau
        up = (cand >= twice) & (np.fmod(np.full_like(tau, float(t_next)), twice) == 0)
        dt_new[...] = np.where(up, twice, np.minimum(cand, tau))
        time_new[...] = t_next
