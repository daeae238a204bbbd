"""Synthetic initial conditions for the benchmarks and parity tests.

``make_plummer`` is a vectorised sampler of the Plummer (1911) sphere in N-body units
(G = M = 1, E = -1/4), equal masses ``1/n``, per-particle ``eps2 = eps^2 / 2`` so that the
pairwise sum ``ie2 + je2`` is ``eps^2`` (the reference's convention,
``tupan/ics/plummer.py:62-63``), default ``eps = 4/n`` (``tupan/tests/test_plummer.py:16-24``).
The reference sampler (``ics/plummer.py:65-133``) needs an O(N^2) potential and a Python
rejection loop per particle, which is impractical at N = 2^20; this one uses the analytic
potential and the standard Aarseth-Henon-Wielen (1974) rejection in bulk.  It produces the
same distribution, not the same sample.
"""
import numpy as np

from .particles import ParticleSystem


def _unit_vectors(rng, n):
    z = rng.uniform(-1.0, 1.0, n)
    ph = rng.uniform(0.0, 2.0 * np.pi, n)
    s = np.sqrt(1.0 - z * z)
    return s * np.cos(ph), s * np.sin(ph), z


def make_plummer(n, eps=None, seed=1, dtype=np.float64, mfrac=0.999):
    n = max(int(n), 2)
    rng = np.random.default_rng(seed)
    if eps is None:
        eps = 4.0 / n
    # radii from the cumulative mass profile, stratified like ics/plummer.py:67-68
    strata = rng.permutation(n)
    mr = (strata + rng.random(n)) * mfrac / n
    r = 1.0 / np.sqrt(mr ** (-2.0 / 3.0) - 1.0)
    ux, uy, uz = _unit_vectors(rng, n)
    rx, ry, rz = r * ux, r * uy, r * uz
    # speeds: q = v/v_esc with density g(q) = q^2 (1-q^2)^(7/2), rejection in bulk
    q = np.empty(n)
    todo = np.arange(n)
    while todo.size:
        x = rng.random(todo.size)
        y = rng.random(todo.size) * 0.1
        ok = y < x * x * (1.0 - x * x) ** 3.5
        q[todo[ok]] = x[ok]
        todo = todo[~ok]
    v = q * np.sqrt(2.0) * (1.0 + r * r) ** (-0.25)
    ux, uy, uz = _unit_vectors(rng, n)
    vx, vy, vz = v * ux, v * uy, v * uz
    # structural -> N-body units (E = -1/4): lengths * 3pi/16, speeds / sqrt(3pi/16)
    sf = 3.0 * np.pi / 16.0
    ps = ParticleSystem(n, np.float64)
    ps.mass[...] = 1.0 / n
    ps.eps2[...] = eps * eps / 2.0
    for name, a in (("rx", rx), ("ry", ry), ("rz", rz)):
        getattr(ps, name)[...] = a * sf
    for name, a in (("vx", vx), ("vy", vy), ("vz", vz)):
        getattr(ps, name)[...] = a / np.sqrt(sf)
    # centre of mass to the origin
    for name in ("rx", "ry", "rz", "vx", "vy", "vz"):
        a = getattr(ps, name)
        a -= a.mean()
    return ps if np.dtype(dtype) == np.float64 else ps.astype(dtype)


def make_uniform(n, seed=0, dtype=np.float64, eps2=0.0):
    """The adversarial set of ``tupan/tests/test_extensions.py:28-36``: mass U(0,1),
    eps2 = 0 (exercises the r2 > 0 mask), positions and velocities U(0,10); seeded."""
    rng = np.random.default_rng(seed)
    ps = ParticleSystem(n, np.float64)
    ps.mass[...] = rng.random(n)
    ps.eps2[...] = eps2
    for name in ("rx", "ry", "rz", "vx", "vy", "vz"):
        getattr(ps, name)[...] = rng.random(n) * 10
    return ps if np.dtype(dtype) == np.float64 else ps.astype(dtype)


def make_binary_rich(n, relative_size=1.0e-3, m_ratio=1.0, ecc=0.5, seed=1, dtype=np.float64, eps2=0.0,
                     orient="random"):
    """Binary-rich Plummer sphere (BASELINE.json configs[4]): every star of an n/2-body Plummer
    model is replaced by a binary of the same total mass whose radius of gyration is
    ``relative_size`` times the parent's -- what the reference builds with
    ``make_hierarchy(make_plummer(n/2, ...), relative_size, make_binary, m1, m2, a, e)``
    (``tupan/ics/hierarchy.py:12-29``, ``ics/fewbody.py:13-50``), vectorised.  Binaries start
    at apocentre.  ``orient="reference"`` keeps the reference's layout (separation along x,
    velocities along y for every binary); ``"random"`` draws an isotropic orientation per
    binary.  Pair b is particles (2b, 2b+1).  Default ``eps2 = 0``: with softening, tight
    binaries drive the reference's Kepler energy check into unbounded sub-stepping
    (Sakura / Kepler)."""
    nb = max(int(n) // 2, 1)
    parent = make_plummer(nb, seed=seed)
    rng = np.random.default_rng(seed + 7919)
    M = parent.mass
    m1 = M * m_ratio / (1.0 + m_ratio)
    m2 = M - m1
    # radius of gyration of the parent (particles/body.py:387-402)
    rr = parent.rx ** 2 + parent.ry ** 2 + parent.rz ** 2
    size = relative_size * np.sqrt(np.sum(M * rr) / np.sum(M))
    # a binary at apocentre with separation d has gyration radius d sqrt(m1 m2)/M
    d = size * M / np.sqrt(m1 * m2)
    a = d / (1.0 + ecc)
    v = np.sqrt(M / a * (1.0 - ecc) / (1.0 + ecc))
    if orient == "reference":
        ex = np.stack([np.ones(nb), np.zeros(nb), np.zeros(nb)])
        ey = np.stack([np.zeros(nb), np.ones(nb), np.zeros(nb)])
    else:
        ex = np.stack(_unit_vectors(rng, nb))
        t = np.stack(_unit_vectors(rng, nb))
        ey = np.cross(ex.T, t.T).T
        ey /= np.sqrt((ey ** 2).sum(0))
    ps = ParticleSystem(2 * nb, np.float64)
    f1, f2 = m2 / M, -m1 / M
    for k, (mk, f) in enumerate(((m1, f1), (m2, f2))):
        ps.mass[k::2] = mk
        for c, name in enumerate(("rx", "ry", "rz")):
            getattr(ps, name)[k::2] = getattr(parent, name) + f * d * ex[c]
        for c, name in enumerate(("vx", "vy", "vz")):
            getattr(ps, name)[k::2] = getattr(parent, name) + f * v * ey[c]
    ps.eps2[...] = eps2
    return ps if np.dtype(dtype) == np.float64 else ps.astype(dtype)
