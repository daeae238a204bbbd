"""Individual block time-steps (tupan_b200.block.BlockHermite) driven on
numpy arrays with the oracle kernels (oracle/block_ops.py).  The reference has no such
integrator, so the checks are physical: synchronisation, energy conservation, agreement with
the reference-pinned shared-step Hermite (oracle/integrators.py) to truncation error, and
fewer particle steps than the shared scheme needs."""
import numpy as np
import pytest

from oracle import integrators as oi
from oracle.block_ops import OracleOps
from tupan_b200 import ics
from tupan_b200.block import BlockHermite

S8 = ("mass", "rx", "ry", "rz", "eps2", "vx", "vy", "vz")


def energy(b):
    ke, pe = b.energies()
    return ke + pe


@pytest.mark.parametrize("order", (4, 6))
def test_block_steps_are_synchronous_and_conserve_energy(order):
    ps = ics.make_plummer(96, seed=5)
    b = BlockHermite(1.0 / 32, ps, order=order, dt_max=2.0 ** -4, ops=OracleOps())
    e0 = energy(b)
    b.evolve(0.25)
    out = b.download(ics.make_plummer(96, seed=5))
    assert np.all(out.time == 0.25), "every particle is synchronous at a multiple of dt_max"
    dts = np.unique(out.tstep)
    assert np.all(np.log2(dts) == np.round(np.log2(dts))) and dts.max() <= 2.0 ** -4
    assert len(dts) > 1, "a Plummer sphere has more than one step level"
    assert abs((energy(b) - e0) / e0) < (2e-6 if order == 4 else 2e-7)
    # the point of the scheme: far fewer particle updates than block steps x N
    assert b.particle_steps < 0.75 * b.block_steps * b.n


def test_agrees_with_the_shared_step_hermite_to_truncation_error():
    n, eta, t_end = 64, 1.0 / 64, 0.125
    ps = ics.make_plummer(n, seed=9)
    b = BlockHermite(eta, ps, order=4, dt_max=2.0 ** -3, ops=OracleOps())
    b.evolve(t_end)
    out = b.download(ics.make_plummer(n, seed=9))
    ref, _ = oi.evolve({k: getattr(ps, k).copy() for k in S8}, "float64", "ahermite4", eta, t_end)
    for k in ("rx", "ry", "rz", "vx", "vy", "vz"):
        a, r = getattr(out, k), getattr(ref, k)
        assert np.max(np.abs(a - r)) / np.max(np.abs(r)) < 2e-6, k


def test_equal_steps_reduce_to_one_block():
    # light, slow, well separated particles: the criterion exceeds dt_max for all of them, so
    # everybody takes dt_max and each block step advances all
    ps = ics.make_uniform(4, seed=1, eps2=1e-2)
    ps.mass[...] = 1e-3
    for k in ("vx", "vy", "vz"):
        getattr(ps, k)[...] *= 1e-3
    b = BlockHermite(1.0 / 16, ps, order=4, dt_max=2.0 ** -6, ops=OracleOps())
    n0 = b.step()
    assert n0 == 4 and b.t == 2.0 ** -6
