"""Pin oracle/integrators.py (the CPU restatement of the reference's integrators) against
tests/golden/integrators_fp{64,32}.npz, which the UNMODIFIED reference produced
(tests/golden/make_golden_integrators.py).  Kernels underneath: oracle/tupan_oracle.c, which
tests/test_oracle.py pins bit-for-bit against the reference C backend -- so the integrated
states must be bit-identical too."""
import numpy as np
import pytest

from golden_util import load_integrator_cases
from oracle import integrators as oi

OUT = ("rx", "ry", "rz", "vx", "vy", "vz", "time", "tstep", "nstep")


@pytest.mark.parametrize("prec", ("float64", "float32"))
def test_oracle_integrators_reproduce_reference_bit_for_bit(prec):
    cases = load_integrator_cases(prec)
    assert len(cases) >= 28
    for name, (ins, outs, meta) in sorted(cases.items()):
        method = name.rsplit("_n", 1)[0]
        eta, t_end, steps, t_final = meta[:4]
        ps, nsteps = oi.evolve(ins, prec, method, eta, t_end)
        assert nsteps == int(steps), (name, nsteps, steps)
        assert float(ps.clock[0]) == t_final, (name, ps.clock[0], t_final)
        order = np.argsort(ps.a["id"])               # hierarchical SIA reorders (join appends)
        ref_order = np.argsort(outs["id"])
        for k in OUT:
            got, ref = ps.a[k][order], outs[k][ref_order]
            assert got.dtype == ref.dtype, (name, k)
            assert np.array_equal(got, ref), (name, k, np.max(np.abs(got.astype(float) - ref.astype(float))))
        ke0, pe0, ke1, pe1 = meta[4:8]
        assert ps.kinetic_energy == ke1 and ps.potential_energy == pe1, name


def test_base_tstep_and_block_quantisation():
    # integrator/__init__.py:48-78
    assert oi.get_base_tstep(0.0, 1.0, 1.0 / 64) == 1.0 / 64
    assert oi.get_base_tstep(0.99, 1.0, 1.0 / 64) == pytest.approx(0.01)
    assert oi.get_base_tstep(0.0, 1.0, -1.0 / 64) == -1.0 / 64
    assert oi.get_min_block_tstep(0.3, 0.0, 1.0) == 0.25
    assert oi.get_min_block_tstep(0.25, 0.0, 1.0) == 0.125
    assert oi.get_min_block_tstep(0.3, 0.125, 1.0) == 0.125      # commensurate with t_curr
    assert oi.get_min_block_tstep(0.3, 0.0, 1.0 / 64) == 1.0 / 64  # never above the base step
    assert oi.get_min_block_tstep(0.3, 0.0, -1.0) == -0.25


def test_operator_sequences():
    # SIA43.dkd (sia.py:441-453): d0 k0 d1 k1 d1 k0 d0; SIA22 (sia.py:376-386): d0 k0 d1 k0 d0
    A, B = oi.SIA_COEFS["sia43"]
    assert [c for _, c in oi.palindrome(B, A)] == [B[0], A[0], B[1], A[1], B[1], A[0], B[0]]
    A, B = oi.SIA_COEFS["sia22"]
    assert [c for _, c in oi.palindrome(B, A)] == [B[0], A[0], B[1], A[0], B[0]]
    for name, (A, B) in oi.SIA_COEFS.items():       # consistency: drift and kick weights sum to 1
        seq = oi.palindrome(B, A)
        assert abs(sum(c for w, c in seq if w == 0) - 1) < 1e-14, name
        assert abs(sum(c for w, c in seq if w == 1) - 1) < 1e-14, name


def test_oracle_post_newtonian_sia_reproduces_reference_bit_for_bit():
    """kick_pn / drift_pn and the PN bookkeeping arrays (sia.py:90-159, body.py:471-527)."""
    cases = load_integrator_cases("float64", "integrators_pn")
    assert len(cases) >= 5
    for name, (ins, outs, meta) in sorted(cases.items()):
        method = name.split("_n", 1)[0]
        eta, t_end, steps, t_final = meta[:4]
        pn = (int(meta[8]), float(meta[9]))
        ps, nsteps = oi.evolve(ins, "float64", method, eta, t_end, pn=pn)
        assert nsteps == int(steps) and float(ps.clock[0]) == t_final, name
        for k in outs:
            if k == "id":
                continue
            assert np.array_equal(ps.a[k], outs[k]), (name, k)
        assert ps.kinetic_energy == meta[6] and ps.potential_energy == meta[7], name
