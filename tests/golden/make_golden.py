"""Generate tests/golden/*.npz by running the UNMODIFIED reference (ggf84/tupan) Python stack.

Run in the build container only (needs /root/reference, gcc, cffi, scipy):

    python tests/golden/make_golden.py            # fp64 vectors
    python tests/golden/make_golden.py --use_sp   # fp32 vectors (the reference reads the
                                                  # precision flag from sys.argv at import,
                                                  # tupan/lib/utils/ctype.py:13)

Everything goes through the reference's public path: `tupan.ics.plummer.make_plummer` /
`tupan.ics.fewbody.make_binary` for inputs, `tupan.lib.extensions.<kernel>.calc(ips, jps, ...)`
(cffi C backend) for outputs.  The files hold inputs, scalars and outputs so the tests need
neither the reference nor this script.
"""
import os
import sys
import tempfile

os.environ["HOME"] = tempfile.mkdtemp(prefix="tupan_home_")   # ~/.tupan/cffi-cache-* must be writable
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
from tupan.lib import extensions as ext  # noqa: E402
from tupan.lib.utils import ctype  # noqa: E402
from tupan.ics.plummer import make_plummer  # noqa: E402
from tupan.particles.allparticles import ParticleSystem  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TAG = "fp32" if ctype.use_sp else "fp64"
S8 = ("mass", "rx", "ry", "rz", "eps2", "vx", "vy", "vz")


def snapshot(ps, extra=()):
    return {k: np.array(getattr(ps, k)) for k in S8 + tuple(extra)}


def uniform_system(n, seed):
    """tupan/tests/test_extensions.py:28-36 with a seed: eps2 = 0 exercises the r2>0 mask."""
    rng = np.random.RandomState(seed)
    ps = ParticleSystem(n - n // 2, n // 2)
    ps.mass[...] = rng.random_sample((ps.n,))
    ps.eps2[...] = 0
    ps.rx[...], ps.ry[...], ps.rz[...] = rng.random_sample((ps.n, 3)).T * 10
    ps.vx[...], ps.vy[...], ps.vz[...] = rng.random_sample((ps.n, 3)).T * 10
    return ps


def binaries_system(nb, seed):
    """Tight, mildly eccentric binaries scattered in a box: drives sakura into its Kepler
    branch (r2 <= (64 m / v2)^2, sakura_kernel_common.h:111-113)."""
    rng = np.random.RandomState(seed)
    ps = ParticleSystem(2 * nb)
    for b in range(nb):
        m1, m2 = rng.uniform(0.5, 1.5, 2) / nb
        a = rng.uniform(1e-3, 1e-2)
        e = rng.uniform(0.0, 0.7)
        m = m1 + m2
        r = a * (1 + e)                                   # apocentre
        v = np.sqrt(m * (1 - e) / (a * (1 + e)))
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        t = np.cross(d, rng.normal(size=3)); t /= np.linalg.norm(t)
        cm = rng.uniform(-1, 1, 3)
        vcm = rng.normal(size=3) * 0.3
        for k, (mk, s) in enumerate(((m1, m2 / m), (m2, -m1 / m))):
            i = 2 * b + k
            ps.mass[i] = mk
            ps.rx[i], ps.ry[i], ps.rz[i] = cm + s * r * d
            ps.vx[i], ps.vy[i], ps.vz[i] = vcm + s * v * t
    ps.eps2[...] = 0
    return ps


def run_all(ps, nj_list, tag):
    out = {"inputs": snapshot(ps)}
    n = ps.n
    eta, dt, c = 1.0 / 64, 1.0 / 64, 128.0
    ext.clight.clight = c
    for nj in nj_list:
        jps = ps[:nj]
        for (ips, jps_, key) in ((ps, jps, "i%d_j%d" % (n, nj)), (jps, ps, "i%d_j%d" % (nj, n))):
            if key in out:
                continue
            res = {}
            res["phi_kernel"] = [np.array(a) for a in ext.phi.calc(ips, jps_)]
            res["acc_kernel"] = [np.array(a) for a in ext.acc.calc(ips, jps_)]
            res["acc_jerk_kernel"] = [np.array(a) for a in ext.acc_jerk.calc(ips, jps_)]
            res["tstep_kernel"] = [np.array(a) for a in ext.tstep.calc(ips, jps_, eta)]
            for order in (2, 4, 5, 6, 7):
                ext.clight.pn_order = order
                res["pnacc_kernel_o%d" % order] = [np.array(a) for a in ext.pnacc.calc(ips, jps_)]
            res["nreg_Xkernel"] = [np.array(a) for a in ext.nreg_x.calc(ips, jps_, dt)]
            for flag in (-2, -1, 1, 2, 0):
                res["sakura_kernel_f%d" % flag] = [np.array(a) for a in ext.sakura.calc(ips, jps_, dt, flag)]
            out[key] = res
    # kernels that need a/j of BOTH sides: full system only (acc_jerk of ps on ps first)
    ext.acc_jerk.calc(ps, ps)
    out["inputs"].update({k: np.array(getattr(ps, k)) for k in ("ax", "ay", "az", "jx", "jy", "jz")})
    key = "i%d_j%d" % (n, n)
    out[key]["snap_crackle_kernel"] = [np.array(a) for a in ext.snap_crackle.calc(ps, ps)]
    out[key]["nreg_Vkernel"] = [np.array(a) for a in ext.nreg_v.calc(ps, ps, dt)]
    flat = {}
    for k, v in out["inputs"].items():
        flat["in/" + k] = v
    for shape, res in out.items():
        if shape == "inputs":
            continue
        for kname, arrs in res.items():
            for q, a in enumerate(arrs):
                flat["%s/%s/%d" % (shape, kname, q)] = a
    flat["scalars"] = np.array([eta, dt, c])
    path = os.path.join(HERE, "%s_%s.npz" % (tag, TAG))
    np.savez_compressed(path, **flat)
    print(path, len(flat), "arrays")


def kepler_cases():
    rows = []
    for seed in range(6):
        ps = binaries_system(1, 100 + seed)
        if seed >= 3:
            ps.eps2[...] = 1e-6        # softened: exercises the energy-check path
        for dt in (1.0 / 64, 0.37, -0.05):
            before = snapshot(ps)
            q = ps.copy()
            ext.kepler.calc(q, q, dt)
            rows.append((before, dt, snapshot(q)))
    flat = {}
    for n, (b, dt, a) in enumerate(rows):
        for k, v in b.items():
            flat["%d/in/%s" % (n, k)] = v
        for k in ("rx", "ry", "rz", "vx", "vy", "vz"):
            flat["%d/out/%s" % (n, k)] = a[k]
        flat["%d/dt" % n] = np.array(dt)
    path = os.path.join(HERE, "kepler_%s.npz" % TAG)
    np.savez_compressed(path, **flat)
    print(path, len(rows), "cases")


if __name__ == "__main__":
    run_all(make_plummer(48, 4.0 / 48, ("equalmass",), seed=1), (1, 17, 48), "plummer48")
    run_all(uniform_system(40, 7), (1, 2, 23, 40), "uniform40")
    run_all(binaries_system(12, 3), (24,), "binaries24")
    kepler_cases()
