"""GPU parity at the BENCHED configuration and launch shapes ("no parity at the
benched configuration"): BASELINE.json's headline is acc_jerk fp64 at N = 2^20, which the launch
plan runs as a 2-D grid of i-blocks x j-chunks (raw accumulators in a workspace, then
finalize_kernel) -- a shape the small-N tests only reach with 2-5 chunks.

* sampled parity at N = 2^20 with the automatic plan (the oracle on 256 i-particles x the full
  j-set, tolerance 1e-12 as everywhere, test_parity_gpu.py);
* forced plans with 32 and 64 j-chunks at small N, every Newtonian kernel;
* rectangular ni = 65536 x nj = 4194304 (the i-sharded shape of an 8-GPU run at N = 4M, and the
  shape block time-steps produce);
* size-independent properties at full size: Newton's third law for the jerk and the
  acceleration (sum_i m_i a_i = 0, sum_i m_i j_i = 0)."""
import ctypes
import os

import numpy as np
import pytest

import oracle
from util import KERNELS, S8, as_dict, cuda_lib, cuda_run, rel_err, run
from tupan_b200 import backend, ics

pytestmark = pytest.mark.gpu
CORES = os.cpu_count() or 1


@pytest.fixture(autouse=True)
def _auto_plan():
    yield
    for prec in ("float64", "float32"):
        backend.load(prec).tupan_cuda_force_plan(-1, 0, 1)


def reference_lib(prec="float64"):
    """The unmodified reference C backend when it was built here (oracle/_ref), else its pinned
    restatement (bit-identical to it, tests/test_oracle.py)."""
    return oracle.load("ref" if oracle.have("ref", prec) else "oracle", prec)


def oracle_sample(name, data, idx, scalars=None, prec="float64"):
    attrs, default, _ = KERNELS[name]
    scalars = default if scalars is None else scalars
    n = len(data["mass"])
    ia = [np.ascontiguousarray(data[a][idx]) for a in attrs]
    ja = [np.ascontiguousarray(data[a]) for a in attrs]
    outs = [np.zeros(len(idx), np.dtype(prec)) for _ in range(oracle.n_outputs(name))]
    oracle.call_threaded(reference_lib(prec), name, prec, CORES, *([len(idx)] + ia + [n] + ja + list(scalars) + outs))
    return outs


def last_plan(lib):
    p = [ctypes.c_int() for _ in range(3)]
    lib.tupan_cuda_last_plan(*[ctypes.byref(x) for x in p])
    return tuple(x.value for x in p)


def test_acc_jerk_sampled_parity_at_n_2_20():
    n = 1 << 20
    ps = ics.make_plummer(n, seed=1)                       # the bench's system
    data = as_dict(ps, "float64")
    lib = cuda_lib("float64")
    got = cuda_run("acc_jerk_kernel", "float64", data, data)
    lane_split, _, jg = last_plan(lib)
    assert lane_split in (0, 2) and jg >= 8, (lane_split, jg)   # the chunked throughput shape the bench times
    rng = np.random.default_rng(20)
    idx = np.unique(np.concatenate([np.linspace(0, n - 1, 128).astype(np.int64), rng.integers(0, n, 128)]))
    ref = oracle_sample("acc_jerk_kernel", data, idx)
    e = rel_err("acc_jerk_kernel", [g[idx] for g in got], ref)
    assert e <= 1e-12, e
    # Newton's third law at full size: the mass-weighted sums vanish up to rounding
    m = data["mass"]
    for g in got:
        assert abs(np.sum(m * g)) <= 1e-11 * np.sum(np.abs(m * g))


@pytest.mark.parametrize("shape,jg", ((0, 32), (0, 64), (2, 32)))
def test_forced_many_chunk_plans(shape, jg):
    """jg = 32 / 64 chunks over blockIdx.y at a size where every chunk is 1-3 tiles, ragged last
    tile, ni not a multiple of the i-block (768 or, shape 2, 512 particles for acc_jerk)."""
    n = 128 * 67 + 19
    ps = ics.make_plummer(n, seed=3)
    data = as_dict(ps, "float64")
    lib = cuda_lib("float64")
    olib = oracle.load("oracle", "float64")
    rng = np.random.default_rng(jg)
    idx = np.sort(rng.choice(n, 192, replace=False))
    for name in ("acc_jerk_kernel", "acc_kernel", "phi_kernel", "tstep_kernel", "snap_crackle_kernel",
                 "nreg_Xkernel", "nreg_Vkernel", "pnacc_kernel"):
        lib.tupan_cuda_force_plan(shape, 0, jg)
        got = cuda_run(name, "float64", data, data)
        assert last_plan(lib)[2] == jg
        ref = oracle_sample(name, data, idx)
        e = rel_err(name, [g[idx] for g in got], ref)
        assert e <= 1e-12, (name, jg, e)
        lib.tupan_cuda_force_plan(-1, 0, 1)
        auto = cuda_run(name, "float64", data, data)
        assert rel_err(name, got, auto) <= 1e-13, (name, jg)    # same pairs, other summation order


def test_rectangular_65536_x_4194304():
    nj, ni = 1 << 22, 1 << 16
    ps = ics.make_plummer(nj, seed=2)
    J = as_dict(ps, "float64")
    rng = np.random.default_rng(4)
    pick = np.sort(rng.choice(nj, ni, replace=False))
    I = {k: np.ascontiguousarray(v[pick]) for k, v in J.items()}
    got = cuda_run("acc_jerk_kernel", "float64", I, J)
    sub = np.sort(rng.choice(ni, 64, replace=False))
    ia = [np.ascontiguousarray(I[a][sub]) for a in S8]
    ja = [J[a] for a in S8]
    ref = [np.zeros(len(sub)) for _ in range(6)]
    oracle.call_threaded(reference_lib(), "acc_jerk_kernel", "float64", CORES, *([len(sub)] + ia + [nj] + ja + ref))
    e = rel_err("acc_jerk_kernel", [g[sub] for g in got], ref)
    assert e <= 1e-12, e


@pytest.mark.parametrize("shape", (0, 2))
@pytest.mark.parametrize("eps2_case", ("zero", "positive", "mixed"))
def test_grouped_kernel_mask_cases(shape, eps2_case):
    """The grouped acc_jerk kernel folds e2_i into the r2 chain and tests the mask (the reference's
    `r2 > 0`, acc_jerk_kernel_common.h:36) once per group of pairs, re-forming r2 only in groups where a
    pair may be masked.  Every way into and around that rare branch, against the reference:
    the pair of a particle with itself; coincident distinct particles with different velocities (masked
    even when softened: their jerk term m (e2)^-3/2 v must NOT appear); pairs closer than 2^-10 softening
    lengths (the branch is taken, the pair is NOT masked); and all of it with e2 = 0, e2 > 0 and mixed."""
    n = 6 * 768 + 37                                 # whole and ragged i-blocks of both shapes
    ps = ics.make_plummer(n, seed=13)
    rng = np.random.default_rng(5)
    eps = np.sqrt(2 * ps.eps2.max())                 # softening length of a pair
    if eps2_case == "zero":
        ps.eps2[...] = 0
    elif eps2_case == "mixed":
        ps.eps2[rng.random(n) < 0.5] = 0
    dup = rng.choice(n // 2, 300, replace=False)     # coincident with a partner in the other half, own velocity
    for a in ("rx", "ry", "rz"):
        getattr(ps, a)[dup + n // 2] = getattr(ps, a)[dup]
    near = np.setdiff1d(np.arange(n // 2), dup)[:300]  # a partner 1e-5 softening lengths away (not masked)
    for a in ("rx", "ry", "rz"):
        getattr(ps, a)[near + n // 2] = getattr(ps, a)[near] + 1e-5 * eps * rng.standard_normal(len(near))
    data = as_dict(ps, "float64")
    lib = cuda_lib("float64")
    idx = np.unique(np.concatenate([dup, dup + n // 2, near, near + n // 2, rng.integers(0, n, 200)]))
    # acc, phi, tstep and nreg_X run the same scheme in their grouped forms (their Op's group_phase1)
    for name in ("acc_jerk_kernel", "acc_kernel", "phi_kernel", "tstep_kernel", "nreg_Xkernel"):
        ref = oracle_sample(name, data, idx)
        assert all(np.all(np.isfinite(r)) for r in ref)
        for jg in (1, 3):
            lib.tupan_cuda_force_plan(shape, 0, jg)
            got = cuda_run(name, "float64", data, data)
            assert last_plan(lib) == (shape, 0, jg)
            assert all(np.all(np.isfinite(g)) for g in got)
            # per particle, not on the scale of the array: a jerk term that should have been masked is
            # visible on the particle it belongs to
            for lo in range(0, len(ref), 3):
                g = np.stack([x[idx] for x in got[lo:lo + 3]])
                r = np.stack(ref[lo:lo + 3])
                err = np.sqrt(((g - r) ** 2).sum(0)) / np.sqrt((r ** 2).sum(0))
                assert err.max() <= 1e-11, (name, shape, eps2_case, jg, lo, err.max(), idx[err.argmax()])
