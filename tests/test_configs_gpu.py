"""BASELINE.json `configs` as GPU parity cases (configs[3], the throughput sweep, is
tools/sweep.py + bench.py; configs[0] is in test_integrator_gpu.py).

configs[1]  Plummer N=65536, shared-time-step leapfrog (sia21s.dkd: acc + phi), fp64 and fp32,
            energy error and sampled final states against a golden run on the reference's C
            backend (tests/golden/make_golden_config2.py).
configs[2]  Plummer N=262144, Hermite6 (acc_jerk + snap_crackle + tstep): one adaptive step on one
            GPU here, forces of a random i-sample against the oracle, block step against the
            oracle's; the 8-GPU i-sharded run is tools/run_integration.py under torchrun
            (test_sharded_gpu.py covers the sharded integrator at world size 2).
configs[4]  binary-rich Plummer N=16384: pnacc orders 2,4,5,6,7 (clight = 128), sakura flags
            -2,-1,1,2 at dt = 1/64 and 1/1024 on an i-sample against the oracle, and the
            batched Kepler solver on all 8192 binaries.
Tolerances as stated in test_parity_gpu.py (1e-12 fp64 for phi...pnacc, 1e-10 for the solvers)."""
import ctypes
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import integrators as oi
from util import S8, as_dict, cuda_lib, cuda_run, pn_scalars, rel_err, run, state_floors
from tupan_b200 import backend, device, ics
from tupan_b200.integrator import Integrator

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VEC = ("rx", "ry", "rz", "vx", "vy", "vz")


def relmax(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


@pytest.mark.parametrize("prec", ("float64", "float32"))
def test_config2_leapfrog_n65536_energy_error(prec):
    z = np.load(os.path.join(GOLDEN, "config2_leapfrog_n65536.npz"))
    eta, t_end, steps_ref, t_ref, ke0r, pe0r, ke1r, pe1r = z[prec + "/meta"]
    n = 65536
    ps = ics.make_plummer(n, seed=1, dtype=prec)
    it = Integrator(eta, 0.0, ps, method="sia21s.dkd")
    ke0, pe0 = it.energies()
    steps = it.evolve(t_end, check_every=4)
    ke1, pe1 = it.energies()
    assert steps == int(steps_ref) and it.time == t_ref
    # energies of the initial state: same inputs, only the summation order differs
    etol = 1e-12 if prec == "float64" else 2e-6
    assert abs(ke0 / ke0r - 1) < etol and abs(pe0 / pe0r - 1) < etol
    eerr = ((ke1 + pe1) - (ke0 + pe0)) / (-pe1)
    eerr_ref = ((ke1r + pe1r) - (ke0r + pe0r)) / (-pe1r)
    assert abs(eerr - eerr_ref) <= (1e-11 if prec == "float64" else 1e-6), (eerr, eerr_ref)
    out = it.particle_system
    idx = z[prec + "/idx"]
    for k in VEC:
        e = relmax(getattr(out, k)[idx], z["%s/out/%s" % (prec, k)])
        assert e <= (1e-10 if prec == "float64" else 2e-4), (k, e)


def test_config3_hermite6_n262144_one_step():
    n, eta = 262144, 1.0 / 64
    ps = ics.make_plummer(n, seed=1)
    data = as_dict(ps, "float64")
    olib = oracle.load("oracle", "float64")
    rng = np.random.default_rng(11)
    idx = np.sort(rng.choice(n, 256, replace=False))
    it = Integrator(eta, 0.0, ps, method="ahermite6")
    # forces of the initial state, exactly as the step computes them
    it.force("tstep_kernel", ("_tstep", "tstepij"), (eta,))
    it.force("acc_jerk_kernel", ("ax", "ay", "az", "jx", "jy", "jz"))
    it.force("snap_crackle_kernel", ("sx", "sy", "sz", "cx", "cy", "cz"))
    st = {k: it.st[k].cpu().numpy() for k in ("_tstep", "tstepij", "ax", "ay", "az", "jx", "jy", "jz",
                                              "sx", "sy", "sz", "cx", "cy", "cz")}
    sample = {k: np.ascontiguousarray(v[idx]) for k, v in data.items()}
    ref = run(olib, "acc_jerk_kernel", "float64", sample, data)
    got = [st[k][idx] for k in ("ax", "ay", "az", "jx", "jy", "jz")]
    assert rel_err("acc_jerk_kernel", got, ref) <= 1e-12
    ref = run(olib, "tstep_kernel", "float64", sample, data, (eta,))
    assert rel_err("tstep_kernel", [st["_tstep"][idx], st["tstepij"][idx]], ref) <= 1e-12
    # snap_crackle consumes a, j of ALL particles: feed the oracle the GPU's a, j
    full14 = dict(data)
    for k in ("ax", "ay", "az", "jx", "jy", "jz"):
        full14[k] = st[k]
    sample14 = {k: np.ascontiguousarray(v[idx]) for k, v in full14.items()}
    ref = run(olib, "snap_crackle_kernel", "float64", sample14, full14)
    got = [st[k][idx] for k in ("sx", "sy", "sz", "cx", "cy", "cz")]
    # At this N the crackle of a particle with a close neighbour is a badly conditioned
    # expression (jx - alpha ax - beta vx - gamma rx cancels to a small part of its terms), and
    # two fp64 evaluations that round differently (the reference's C, ours with FMAs) differ by
    # more than 1e-12 there.  Where the stated tolerance is exceeded, the x87 extended-precision
    # build of the restatement decides: our error against it must not exceed 1.5x the
    # reference's own (the criterion test_parity_gpu.py already uses for fp32).
    e = rel_err("snap_crackle_kernel", got, ref)
    if e > 1e-12:
        ld = lambda d: {k: v.astype(np.longdouble) for k, v in d.items()}
        truth = run(oracle.load("oracle", "float128"), "snap_crackle_kernel", "float128", ld(sample14), ld(full14))
        e_cuda, e_ref = rel_err("snap_crackle_kernel", got, truth), rel_err("snap_crackle_kernel", ref, truth)
        assert e_cuda <= 1.5 * e_ref, (e, e_cuda, e_ref)
    # one full adaptive step: block step = the reference's quantisation of the GPU's min tstep
    it.evolve_step(1.0)
    tau = oi.get_min_block_tstep(np.abs(st["_tstep"]).min(), 0.0, oi.get_base_tstep(0.0, 1.0, eta))
    assert it.time == tau and it.nsteps == 1
    out = it.particle_system
    assert np.all(out.tstep == tau) and np.all(out.time == tau) and np.all(out.nstep == 1)
    # size-independent property: total momentum is conserved by the step
    for k in ("vx", "vy", "vz"):
        assert abs(np.sum(out.mass * getattr(out, k))) < 1e-13


def test_config5_binary_rich_pn_sakura_kepler():
    prec = "float64"
    n = 16384
    ps = ics.make_binary_rich(n, seed=1)
    data = as_dict(ps, prec)
    olib = oracle.load("oracle", prec)
    rng = np.random.default_rng(5)
    pairs = np.sort(rng.choice(n // 2, 96, replace=False))
    idx = np.sort(np.concatenate([2 * pairs, 2 * pairs + 1]))          # both members of 96 binaries
    sample = {k: np.ascontiguousarray(v[idx]) for k, v in data.items()}
    for order in (2, 4, 5, 6, 7):
        sc = pn_scalars(order, 128.0)
        got = cuda_run("pnacc_kernel", prec, data, data, sc)
        ref = run(olib, "pnacc_kernel", prec, sample, data, sc)
        e = rel_err("pnacc_kernel", [g[idx] for g in got], ref)
        assert e <= 1e-12, ("pnacc", order, e)
    floors = state_floors(data)
    for dt in (1.0 / 64, 1.0 / 1024):
        for flag in (-2, -1, 1, 2):
            got = cuda_run("sakura_kernel", prec, data, data, (dt, flag))
            ref = run(olib, "sakura_kernel", prec, sample, data, (dt, flag))
            e = rel_err("sakura_kernel", [g[idx] for g in got], ref, floors)
            assert e <= 1e-10, ("sakura", dt, flag, e)
    # every binary through the batched Kepler entry point (SIA leaves with n = 2, sia.py:311-312)
    lib = cuda_lib(prec)
    dt = 1.0 / 64
    d = {k: torch.from_numpy(data[k]).cuda() for k in S8}
    outs = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(6)]
    rc = lib.tupan_cuda_kepler_dev(n // 2, device.ptr_array([d[k] for k in S8]), dt, device.ptr_array(outs),
                                   device.current_stream())
    assert rc == 0
    torch.cuda.synchronize()
    assert lib.tupan_cuda_kepler_limit_hits() == 0
    got = [o.cpu().numpy() for o in outs]
    worst = 0.0
    for b in pairs:
        arrs = [np.ascontiguousarray(data[k][2 * b:2 * b + 2]).copy() for k in S8]
        res = [arrs[1], arrs[2], arrs[3], arrs[5], arrs[6], arrs[7]]
        oracle.call(olib, "kepler_solver_kernel", prec, *(arrs + [dt] + res))
        for lo in (0, 3):
            g = np.stack([got[lo + c][2 * b:2 * b + 2] for c in range(3)])
            r = np.stack(res[lo:lo + 3])
            worst = max(worst, float(np.max(np.sqrt(((g - r) ** 2).sum(0)) / np.sqrt((r ** 2).sum(0)))))
    assert worst <= 1e-10, worst


@pytest.mark.parametrize("method", ("asakura", "sakura", "sia21s.kdk"))
def test_config5_integration_runs_match_the_c_backend(method):
    """BASELINE.json configs[4] as an INTEGRATION: binary-rich Plummer N = 16384,
    Sakura/Kepler pairwise propagation (asakura: 128 adaptive steps; sakura: 2 shared steps) and the
    post-Newtonian SIA (sia21s.kdk, pn_order 7, clight 128; two shared steps), against golden runs on the reference's
    C backend (tests/golden/make_golden_config5.py).  Same number of steps and final clock; relative
    energy error within 1e-10 of the C backend's; sampled binary members within 1e-9 (positions and
    velocities relative to the largest component)."""
    path = os.path.join(GOLDEN, "config5_binary_rich_n16384.npz")
    if not os.path.exists(path):
        pytest.skip("golden runs not generated (tests/golden/make_golden_config5.py)")
    z = np.load(path)
    if method + "/meta" not in z.files:
        pytest.skip("golden run of %s not generated" % method)
    eta, t_end, steps_ref, t_ref, ke0r, pe0r, ke1r, pe1r, pn_order, clight = z[method + "/meta"]
    n = 16384
    ps = ics.make_binary_rich(n, seed=1)
    kw = dict(pn_order=int(pn_order), clight=float(clight)) if pn_order else {}
    it = Integrator(eta, 0.0, ps, method=method, **kw)
    ke0, pe0 = it.energies()
    steps = it.evolve(t_end, check_every=4)
    ke1, pe1 = it.energies()
    assert steps == int(steps_ref) and it.time == t_ref, (steps, steps_ref, it.time, t_ref)
    assert abs(ke0 / ke0r - 1) < 1e-12 and abs(pe0 / pe0r - 1) < 1e-12
    eerr = ((ke1 + pe1) - (ke0 + pe0)) / (-pe1)
    eerr_ref = ((ke1r + pe1r) - (ke0r + pe0r)) / (-pe1r)
    assert abs(eerr - eerr_ref) <= 1e-10, (eerr, eerr_ref)
    out = it.particle_system
    idx = z["idx"]
    for k in VEC:
        e = relmax(getattr(out, k)[idx], z["%s/out/%s" % (method, k)])
        assert e <= 1e-9, (method, k, e)
