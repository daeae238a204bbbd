"""GPU parity tests: the CUDA path, called through the C ABI of include/libtupan_cuda.h (the ten
libtupan.h entry points, host pointers), against the CPU oracle and the committed golden
vectors produced by the reference's own Python stack.

Stated tolerances (per particle, vector-norm relative: ||got - ref||_2 / ||ref||_2 over each
output 3-vector, plain relative for scalar outputs):

    phi, acc, acc_jerk, snap_crackle, tstep, pnacc, nreg_X, nreg_V :  1e-12 fp64, 1e-5 fp32
    sakura, kepler (iterative solver, libm inside)                 :  1e-10 fp64, 1e-3 fp32

The summation order differs from the reference's sequential j loop (tiles, lane split,
j chunks), which is why the comparison is not bit-exact.  Two refinements, both measured on
the B200 (tools/parity_report.py, profiles/parity_r01.txt):

* fp32, N >= 1000: the reference's own sequential fp32 sum is ~1e-5 away from the exact
  result (it grows with N), so "within 1e-5 of the reference" stops being a statement about
  the CUDA kernel.  There the test accepts either 1e-5 against the fp32 reference or -- with
  the fp64 oracle on the same fp32 inputs as the truth -- a CUDA error no larger than 1.5x the
  reference's own error.
* sakura: dr/dv are differences of nearly equal two-body states; their norm can be 1e-9 of
  the operands, so the denominator is floored by the rms position / velocity of the system
  they are added to (util.state_floors).
"""
import ctypes

import numpy as np
import pytest

import oracle
from golden_util import load_kepler, load_set, split_case
from util import KERNELS, S8, as_dict, cuda_lib, cuda_run, pn_scalars, rel_err, run, state_floors
from tupan_b200 import backend, ics

pytestmark = pytest.mark.gpu

PRECS = ("float64", "float32")
TOL = {"float64": 1e-12, "float32": 1e-5}
TOL_SOLVER = {"float64": 1e-10, "float32": 1e-3}
# launch plans (lane_split, js_log2, jg): automatic, i-per-thread, lane split x1/x8/x32 with
# j chunks, i-per-thread with j chunks, the second group shape of the grouped kernels (2; kernels
# without one run shape 0).  Every plan must give the same answer.
PLANS = ((-1, 0, 1), (0, 0, 1), (1, 0, 1), (1, 3, 1), (1, 5, 3), (0, 0, 2), (2, 0, 1), (2, 0, 3))


def tol_for(name, prec):
    return (TOL_SOLVER if name in ("sakura_kernel", "kepler_solver_kernel") else TOL)[prec]


def floors_for(name, I):
    return state_floors(I) if name == "sakura_kernel" else None


def assert_parity(name, prec, got, ref, I, truth=None, what=()):
    """Stated tolerance against the reference; for fp32 optionally the truth-based criterion."""
    fl = floors_for(name, I)
    e = rel_err(name, got, ref, fl)
    if e <= tol_for(name, prec):
        return e
    if prec == "float32" and truth is not None:
        e_cuda, e_ref = rel_err(name, got, truth, fl), rel_err(name, ref, truth, fl)
        assert e_cuda <= 1.5 * e_ref, (name, prec, what, e, e_cuda, e_ref)
        return e
    raise AssertionError((name, prec, what, e))


def scalars_for(kernel, variant, eta, dt, c):
    if kernel == "tstep_kernel":
        return (eta,)
    if kernel == "pnacc_kernel":
        return pn_scalars(variant[1], c)
    if kernel in ("nreg_Xkernel", "nreg_Vkernel"):
        return (dt,)
    if kernel == "sakura_kernel":
        return (dt, variant[1])
    return ()


@pytest.fixture(autouse=True)
def _auto_plan():
    yield
    for prec in PRECS:
        backend.load(prec).tupan_cuda_force_plan(-1, 0, 1)


def test_native_library_is_what_runs():
    """The product path is the in-tree CUDA library: kernels launched are counted by it."""
    lib = cuda_lib("float64")
    before = lib.tupan_cuda_launch_count()
    ps = ics.make_plummer(300, seed=4)
    ps.set_acc_jerk(ps)
    assert lib.tupan_cuda_launch_count() >= before + 2      # pack + pair kernel
    assert lib.tupan_cuda_sm_count() >= 100
    assert np.all(np.isfinite(ps.ax)) and np.all(np.isfinite(ps.jz))


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("setname", ("plummer48", "uniform40", "binaries24"))
def test_golden_vectors_from_reference(setname, prec):
    """Outputs of the unmodified reference (tests/golden/make_golden.py), incl. rectangular
    ni != nj calls in both orientations and eps2 = 0 inputs that exercise the r2 > 0 mask."""
    inputs, shapes, (eta, dt, c) = load_set(setname, prec)
    checked = 0
    for (ni, nj), cases in shapes.items():
        I = {k: v[:ni] for k, v in inputs.items()}
        J = {k: v[:nj] for k, v in inputs.items()}
        for case, ref in cases.items():
            kernel, variant = split_case(case)
            got = cuda_run(kernel, prec, I, J, scalars_for(kernel, variant, eta, dt, c))
            assert_parity(kernel, prec, got, ref, I, None, (setname, case, ni, nj))
            checked += 1
    assert checked >= 17


@pytest.mark.parametrize("prec", PRECS)
def test_kepler_golden_in_place(prec):
    """kepler_solver_kernel: two bodies, outputs alias inputs (extensions.py:642-646).

    All 18 golden cases of the reference must match, the softened ones included: there the
    reference's energy check doubles the number of sub-steps without bound (2^14 .. 2^27 SEQUENTIAL
    sub-steps in these vectors) and so does the two-body entry point (bound 2^27).  The one case
    that needs 2^27 sub-steps (43 s on a host core, minutes on one GPU thread) only runs with
    TUPAN_SLOW_TESTS=1; profiles/r02_kepler_softened.txt keeps its result."""
    import os
    lib = cuda_lib(prec)
    dt_np = np.dtype(prec)
    matched = skipped_slow = 0
    for case, (ins, dt, outs) in enumerate(load_kepler(prec)):
        if prec == "float64" and case == 10 and not os.environ.get("TUPAN_SLOW_TESTS"):
            skipped_slow += 1
            continue
        # The solver removes whole periods from dt (universal_kepler_solver.h:406-412), so an
        # error eps in the period becomes a phase error eps * (number of revolutions): the
        # tolerance is stated per revolution.
        i64 = {k: np.asarray(v, np.float64) for k, v in ins.items()}
        m = i64["mass"].sum()
        r = np.sqrt(sum((i64[k][0] - i64[k][1]) ** 2 for k in ("rx", "ry", "rz")) + i64["eps2"].sum())
        v2 = sum((i64[k][0] - i64[k][1]) ** 2 for k in ("vx", "vy", "vz"))
        alpha = v2 - 2 * m / r
        revs = abs(dt) / (2 * np.pi * m / abs(alpha) ** 1.5) if alpha < 0 else 0.0
        tol = tol_for("kepler_solver_kernel", prec) * max(1.0, revs)
        arrs = [np.ascontiguousarray(ins[a], dt_np).copy() for a in S8]
        res = [arrs[1], arrs[2], arrs[3], arrs[5], arrs[6], arrs[7]]      # in place
        oracle.call(lib, "kepler_solver_kernel", prec, *(arrs + [dt] + res))
        backend.check(lib, "kepler_solver_kernel")                       # no case may be refused
        for lo, names in ((0, ("rx", "ry", "rz")), (3, ("vx", "vy", "vz"))):
            g = np.stack(res[lo:lo + 3]).astype(np.float64)
            r = np.stack([outs[k] for k in names]).astype(np.float64)
            e = np.sqrt(((g - r) ** 2).sum(0)) / np.sqrt((r ** 2).sum(0))
            assert e.max() <= tol, (prec, case, dt, float(ins["eps2"][0]), e.max())
        matched += 1
    assert matched + skipped_slow == 18 and skipped_slow <= 1, (matched, skipped_slow)


@pytest.mark.parametrize("prec", PRECS)
def test_sakura_softened_binaries_take_the_cleanup_launch(prec):
    """Softened tight binaries: pairs whose Kepler sub-stepping needs more than the sweep's 2^12
    sub-steps are handed to the clean-up launch and must still come out as the reference's
    (sakura_kernel_common.h:94-123 -> universal_kepler_solver.h:523-606)."""
    ps = ics.make_binary_rich(48, relative_size=5.0e-3, seed=3, eps2=5.0e-7)
    data = as_dict(ps, prec)
    olib = oracle.load("oracle", prec)
    lib = cuda_lib(prec)
    before = lib.tupan_cuda_kepler_cleanup_pairs()
    floors = state_floors(data)
    for sc in ((1.0 / 64, 1), (1.0 / 64, -2), (-0.05, 2)):
        ref = run(olib, "sakura_kernel", prec, data, data, sc)
        got = cuda_run("sakura_kernel", prec, data, data, sc)
        e = rel_err("sakura_kernel", got, ref, floors)
        assert e <= tol_for("sakura_kernel", prec), (prec, sc, e)
    if prec == "float64":
        assert lib.tupan_cuda_kepler_cleanup_pairs() > before          # the path was exercised


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("name", sorted(KERNELS))
def test_every_plan_matches_oracle(name, prec):
    """Plummer N=1000 (not a multiple of any tile) under every launch plan."""
    data = as_dict(ics.make_plummer(1000, seed=1), prec)
    olib = oracle.load("oracle", prec)
    lib = cuda_lib(prec)
    variants = [None]
    if name == "pnacc_kernel":
        variants = [pn_scalars(k) for k in (0, 1, 2, 3, 4, 5, 6, 7)]
    if name == "sakura_kernel":
        variants = [(1.0 / 64, f) for f in (-2, -1, 1, 2, 0)]
    for sc in variants:
        ref = run(olib, name, prec, data, data, sc)
        truth = None
        if prec == "float32":
            d64 = {k: v.astype(np.float64) for k, v in data.items()}
            truth = run(oracle.load("oracle", "float64"), name, "float64", d64, d64, sc)
        for plan in PLANS:
            lib.tupan_cuda_force_plan(*plan)
            got = cuda_run(name, prec, data, data, sc)
            assert_parity(name, prec, got, ref, data, truth, (sc and sc[:2], plan))


@pytest.mark.parametrize("prec", PRECS)
def test_rectangular_sweep_eps0_both_orientations(prec):
    """The reference's own test shape (tupan/tests/test_extensions.py:39-127): 256 bodies,
    mass U(0,1), eps2 = 0, pos/vel U(0,10); (ips=ps, jps=ps[:jdx]) and transposed."""
    data = as_dict(ics.make_uniform(256, seed=1), prec)
    olib = oracle.load("oracle", prec)
    o64 = oracle.load("oracle", "float64")
    d64 = {k: v.astype(np.float64) for k, v in data.items()}
    for name in sorted(KERNELS):
        sc = None
        if name == "sakura_kernel":
            sc = (1.0 / 64, -1)
        for jdx in (1, 2, 3, 31, 32, 33, 127, 129, 255, 256):
            for (ni, nj) in ((256, jdx), (jdx, 256)):
                ref = run(olib, name, prec, data, data, sc, ni, nj)
                got = cuda_run(name, prec, data, data, sc, ni, nj)
                truth = None
                if prec == "float32":
                    truth = run(o64, name, "float64", d64, d64, sc, ni, nj)
                assert_parity(name, prec, got, ref, data, truth, (ni, nj))


@pytest.mark.parametrize("prec", PRECS)
def test_empty_and_degenerate_shapes(prec):
    data = as_dict(ics.make_plummer(64, seed=3), prec)
    olib = oracle.load("oracle", prec)
    for name in sorted(KERNELS):
        # nj = 0: outputs are the epilogue of empty sums (0, or eta for tstep)
        ref = run(olib, name, prec, data, data, None, 64, 0)
        got = cuda_run(name, prec, data, data, None, 64, 0)
        for g, r in zip(got, ref):
            assert np.array_equal(g, r), (name, "nj=0")
        # ni = 0: nothing is written, nothing fails
        got = cuda_run(name, prec, data, data, None, 0, 64)
        assert all(len(g) == 0 for g in got)
        # single particle on itself: the masked self pair only
        ref = run(olib, name, prec, data, data, None, 1, 1)
        got = cuda_run(name, prec, data, data, None, 1, 1)
        for g, r in zip(got, ref):
            assert np.array_equal(g, r), (name, "1x1")


@pytest.mark.parametrize("prec", PRECS)
def test_coincident_particles_are_masked(prec):
    """Duplicated positions (r2 == 0 for i != j) contribute nothing, as in the reference."""
    ps = ics.make_plummer(128, seed=9)
    for a in ("rx", "ry", "rz"):
        getattr(ps, a)[64:] = getattr(ps, a)[:64]
    ps.eps2[...] = 0
    data = as_dict(ps, prec)
    olib = oracle.load("oracle", prec)
    for name in ("phi_kernel", "acc_kernel", "acc_jerk_kernel", "tstep_kernel", "nreg_Xkernel", "pnacc_kernel"):
        ref = run(olib, name, prec, data, data)
        got = cuda_run(name, prec, data, data)
        assert all(np.all(np.isfinite(g)) for g in got), name
        assert rel_err(name, got, ref) <= TOL[prec], name


def test_fp64_abi_uses_64bit_counts():
    """fp64 library: UINT = unsigned long; fp32: unsigned int (cffi_backend.py:35-36)."""
    assert backend.load("float64").acc_kernel.argtypes[0] is ctypes.c_ulong
    assert backend.load("float32").acc_kernel.argtypes[0] is ctypes.c_uint


def test_sampled_parity_at_65536():
    """BASELINE configs[1] size: oracle on a random i-sample against the full j-set."""
    n = 65536
    ps = ics.make_plummer(n, seed=1)
    data = as_dict(ps, "float64")
    olib = oracle.load("oracle", "float64")
    rng = np.random.default_rng(7)
    idx = np.sort(rng.choice(n, 384, replace=False))
    sample = {k: np.ascontiguousarray(v[idx]) for k, v in data.items()}
    for name in ("acc_jerk_kernel", "acc_kernel", "phi_kernel", "tstep_kernel"):
        got = cuda_run(name, "float64", data, data)
        ref = run(olib, name, "float64", sample, data)
        e = rel_err(name, [g[idx] for g in got], ref)
        assert e <= 1e-12, (name, e)
        if name == "acc_kernel":
            # size-independent property: Newton's third law, sum_i m_i a_i = 0
            m = data["mass"]
            scale = np.abs(m * got[0]).sum()
            for g in got:
                assert abs((m * g).sum()) <= 1e-11 * scale
    # fp32 at the same size
    d32 = as_dict(ps, "float32")
    s32 = {k: np.ascontiguousarray(v[idx]) for k, v in d32.items()}
    got = cuda_run("acc_kernel", "float32", d32, d32)
    ref = run(oracle.load("oracle", "float32"), "acc_kernel", "float32", s32, d32)
    # sequential fp32 summation over 65536 terms in the reference carries ~1e-4 of rounding
    # noise itself; compare both against the fp64 truth instead of against each other
    truth = cuda_run("acc_kernel", "float64", data, data)
    e_cuda = rel_err("acc_kernel", [g[idx] for g in got], [t[idx] for t in truth])
    e_ref = rel_err("acc_kernel", ref, [t[idx] for t in truth])
    assert e_cuda <= max(1e-5, e_ref), (e_cuda, e_ref)


def test_device_resident_api_matches_host_api():
    import torch
    from tupan_b200 import device
    ps = ics.make_plummer(5000, seed=11)
    d = device.to_device(ps)
    out = device.run("acc_jerk_kernel", d, d)
    torch.cuda.synchronize()
    ps.set_acc_jerk(ps)
    for k in ("ax", "ay", "az", "jx", "jy", "jz"):
        assert np.array_equal(out[k].cpu().numpy(), getattr(ps, k)), k
    # fused |tstep| minimum
    out = device.run("tstep_kernel", d, d, (1.0 / 64,))
    lib = cuda_lib("float64")
    mn = torch.empty(1, dtype=torch.float64, device="cuda")
    rc = lib.tupan_cuda_abs_min_dev(5000, ctypes.c_void_p(out["tstep"].data_ptr()),
                                    ctypes.c_void_p(mn.data_ptr()), device.current_stream())
    assert rc == 0
    torch.cuda.synchronize()
    assert mn.item() == out["tstep"].abs().min().item()


def test_i_and_j_aliasing_and_prefix_slices():
    """ips is jps, and jps a prefix view of ips (same host base pointers)."""
    ps = ics.make_plummer(777, seed=5)
    data = as_dict(ps, "float64")
    olib = oracle.load("oracle", "float64")
    ref = run(olib, "acc_jerk_kernel", "float64", data, data, None, 777, 100)
    got = cuda_run("acc_jerk_kernel", "float64", data, data, None, 777, 100)
    assert rel_err("acc_jerk_kernel", got, ref) <= 1e-12
