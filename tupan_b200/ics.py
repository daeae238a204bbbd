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
