"""CPU tests: pin the oracle (oracle/tupan_oracle.c) to the reference.

1. against the committed golden vectors produced by the reference's own Python stack;
2. bit-for-bit against oracle/_ref (the unmodified reference C compiled in place), when that
   library is present (it is built here from /root/reference and travels to the GPU box).
"""
import numpy as np
import pytest

import oracle
from golden_util import load_kepler, load_set, split_case
from util import KERNELS, S8, pn_scalars, rel_err, run
from tupan_b200 import ics

PRECS = ("float64", "float32")
SETS = ("plummer48", "uniform40", "binaries24")


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


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("setname", SETS)
def test_oracle_reproduces_reference_golden(setname, prec):
    inputs, shapes, (eta, dt, c) = load_set(setname, prec)
    lib = oracle.load("oracle", prec)
    n = len(inputs["mass"])
    checked = 0
    for (ni, nj), cases in shapes.items():
        # rectangular cases are prefix slices of the same system (make_golden.py: ps[:nj])
        I = {k: v[:ni] for k, v in inputs.items()}
        J = {k: v[:nj] for k, v in inputs.items()}
        for case, ref in cases.items():
            kernel, variant = split_case(case)
            got = run(lib, kernel, prec, I, J, scalars_for(kernel, variant, eta, dt, c))
            if kernel == "sakura_kernel":
                # libm (cos/sin/cosh/sinh/log) sits inside the Kepler branch
                tol = 1e-13 if prec == "float64" else 1e-6
                assert rel_err(kernel, got, ref) <= tol, (setname, prec, case, ni, nj)
            else:
                for g, r in zip(got, ref):
                    assert np.array_equal(g, r, equal_nan=True), (setname, prec, case, ni, nj)
            checked += 1
    assert checked >= 17 and n > 0


@pytest.mark.parametrize("prec", PRECS)
def test_oracle_kepler_golden(prec):
    lib = oracle.load("oracle", prec)
    dt_np = np.dtype(prec)
    for ins, dt, outs in load_kepler(prec):
        arrs = [np.ascontiguousarray(ins[a], dt_np) for a in S8]
        res = [np.zeros(2, dt_np) for _ in range(6)]
        oracle.call(lib, "kepler_solver_kernel", prec, *(arrs + [dt] + res))
        tol = 1e-13 if prec == "float64" else 1e-5
        for r, k in zip(res, ("rx", "ry", "rz", "vx", "vy", "vz")):
            np.testing.assert_allclose(r, outs[k], rtol=tol, atol=tol * np.abs(outs[k]).max())


@pytest.mark.skipif(not oracle.have("ref", "float64"), reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("eps2", (0.0, 1e-4))
def test_oracle_bit_exact_vs_compiled_reference(prec, eps2):
    from util import as_dict
    ps = as_dict(ics.make_uniform(150, seed=3, eps2=eps2), prec)
    pl = as_dict(ics.make_plummer(128, seed=5), prec)
    o, r = oracle.load("oracle", prec), oracle.load("ref", prec)
    for data in (ps, pl):
        n = len(data["mass"])
        for name in KERNELS:
            variants = [None]
            if name == "pnacc_kernel":
                variants = [pn_scalars(k) for k in (0, 1, 2, 3, 4, 5, 6, 7)]
            if name == "sakura_kernel":
                variants = [(1.0 / 64, f) for f in (-2, -1, 1, 2, 0)]
            for sc in variants:
                for (ni, nj) in ((n, n), (n, 37), (9, n), (n, 0)):
                    a = run(o, name, prec, data, data, sc, ni, nj)
                    b = run(r, name, prec, data, data, sc, ni, nj)
                    for x, y in zip(a, b):
                        assert np.array_equal(x, y, equal_nan=True), (name, prec, eps2, sc, ni, nj)


@pytest.mark.skipif(not oracle.have("ref", "float64"), reason="oracle/_ref not built")
@pytest.mark.parametrize("prec", PRECS)
def test_oracle_kepler_branch_bit_exact(prec):
    """Tight binaries force sakura into the universal Kepler solver; compare with _ref."""
    inputs, _, _ = load_set("binaries24", prec)
    o, r = oracle.load("oracle", prec), oracle.load("ref", prec)
    for dt in (1.0 / 64, 1.0 / 1024, 0.25):
        for flag in (-2, -1, 1, 2):
            a = run(o, "sakura_kernel", prec, inputs, inputs, (dt, flag))
            b = run(r, "sakura_kernel", prec, inputs, inputs, (dt, flag))
            for x, y in zip(a, b):
                assert np.array_equal(x, y, equal_nan=True), (prec, dt, flag)


def test_threaded_i_slices_equal_single_call():
    from util import as_dict
    ps = as_dict(ics.make_plummer(257, seed=2), "float64")
    lib = oracle.load("oracle", "float64")
    ref = run(lib, "acc_jerk_kernel", "float64", ps, ps)
    n = 257
    outs = [np.zeros(n) for _ in range(6)]
    args = [n] + [ps[a] for a in S8] + [n] + [ps[a] for a in S8] + outs
    oracle.call_threaded(lib, "acc_jerk_kernel", "float64", 5, *args)
    for x, y in zip(outs, ref):
        assert np.array_equal(x, y)
