"""CPU checks of the device-resident integrators' host logic (no GPU: no compute)."""
import pytest

from oracle import integrators as oi
from tupan_b200 import ics
from tupan_b200.backend import TupanCudaError
from tupan_b200.integrator import SIA_COEFS, Integrator, operator_sequence


def test_operator_sequences_match_the_pinned_restatement():
    assert SIA_COEFS == oi.SIA_COEFS
    for name, (A, B) in SIA_COEFS.items():
        ours = [w for _, w in operator_sequence(B, A)]
        assert ours == [w for _, w in oi.palindrome(B, A)], name
        assert len(ours) == 2 * (len(A) + len(B)) - 1


def test_method_names_are_the_reference_ones():
    # integrator/hermite.py:291-295, sia.py:969-985, nreg.py:103-104, sakura.py:58-59
    for m in ("hermite2", "ahermite8", "sia21s.dkd", "sia69a.kdk", "sia43h.kdk", "nreg", "anreg", "sakura",
              "asakura"):
        assert m in Integrator.PROVIDED_METHODS
    with pytest.raises(ValueError):
        Integrator(1.0 / 64, 0.0, ics.make_plummer(8), method="rk4", device="cpu")
    with pytest.raises(NotImplementedError):
        Integrator(1.0 / 64, 0.0, ics.make_plummer(8), method="hermite4", device="cpu", pn_order=7, clight=128)
    with pytest.raises(TypeError):          # integrator/__init__.py:27-32
        Integrator(1.0 / 64, 0.0, ics.make_plummer(8), method="sia21s.kdk", device="cpu", pn_order=7)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises((TupanCudaError, RuntimeError, AssertionError)):
        it = Integrator(1.0 / 64, 0.0, ics.make_plummer(8), method="hermite4", device="cuda:0")
        it.evolve_step(1.0)
