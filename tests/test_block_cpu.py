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


def test_step_quantisation_rules():
    # block_quantize (csrc/k_update.cu) restated in oracle/block_ops.py: largest power of two <= the
    # criterion, <= dt_max, at most twice the old step and only when the block time is a multiple of it
    o = OracleOps()
    ts = np.array([0.3, 0.25, 0.2499, 1e-3, 5.0, 0.13, 0.13])
    tau = np.array([0.125, 0.125, 0.125, 0.125, 0.125, 0.03125, 0.03125])
    dt, tm = np.zeros(7), np.zeros(7)
    o.quantize(ts, tau, 0.25, 0.25, dt, tm)          # t = 0.25 is a multiple of 0.25 and of 0.0625
    #   0.3 -> 0.25 = 2 tau, allowed; 0.25 -> 0.25; 0.2499 -> 0.125; 1e-3 -> 2^-10; 5 -> dt_max = 2 tau;
    #   0.13 with tau = 1/32: candidate 1/8 but only one doubling per step -> 1/16
    assert np.array_equal(dt, [0.25, 0.25, 0.125, 2.0 ** -10, 0.25, 0.0625, 0.0625])
    assert np.all(tm == 0.25)
    o.quantize(ts, tau, 0.375, 0.25, dt, tm)         # 0.375 is NOT a multiple of 0.25: no doubling of 1/8
    assert np.array_equal(dt[:5], [0.125, 0.125, 0.125, 2.0 ** -10, 0.125])
    assert np.array_equal(dt[5:], [0.0625, 0.0625])  # 0.375 is a multiple of 1/16


# ---- the i-sharded block step (BASELINE.json configs[2]) with gloo, world sizes 2 and 3 ----------
def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _sharded_worker(rank, world, port, order, n, q):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = BlockHermite(1.0 / 32, ics.make_plummer(n, seed=5), order=order, dt_max=2.0 ** -4, ops=OracleOps())
        assert b.world == world and b.rank == rank
        b.evolve(0.125)
        out = b.download(ics.make_plummer(n, seed=5))
        q.put((rank, b.block_steps, b.particle_steps, b.pairs,
               {k: np.array(getattr(out, k)) for k in ("rx", "ry", "rz", "vx", "vy", "vz", "time", "tstep")}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,order", ((2, 4), (3, 6)))
def test_sharded_block_steps_equal_the_single_process_run(world, order):
    """Replicated state, sharded active set: every rank must end with the same state as one
    process on its own (same kernels on the same pairs; only the batching of the i-set differs,
    and the oracle's j loop is sequential, so the result is bit-identical), and the pair
    interactions must be shared out between the ranks."""
    import torch.multiprocessing as mp
    n = 61                                         # not a multiple of the world size
    ref = BlockHermite(1.0 / 32, ics.make_plummer(n, seed=5), order=order, dt_max=2.0 ** -4, ops=OracleOps())
    ref.evolve(0.125)
    want = ref.download(ics.make_plummer(n, seed=5))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, order, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        got = [q.get(timeout=120) for _ in range(world)]
    finally:
        for p in procs:
            p.join(30)
            if p.is_alive():
                p.terminate()
    assert all(p.exitcode == 0 for p in procs)
    pairs = 0.0
    for rank, bsteps, psteps, prs, state in got:
        assert bsteps == ref.block_steps and psteps == ref.particle_steps
        for k, v in state.items():
            assert np.array_equal(v, getattr(want, k)), (rank, k)
        pairs += prs
        assert prs < 0.75 * ref.pairs              # no rank did (nearly) all the work
    assert pairs == ref.pairs
