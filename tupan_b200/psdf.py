"""Snapshot writer for device-resident state in the reference's PSDF format.

The reference writes snapshots as a YAML stream, one ``--- !Particle`` document per particle
with the keys ``id, m, t, r, v, a`` plus tupan's ``type, eps2, pot[, dt_prev, dt_next]``
(``tupan/io/psdfio.py:25-35,62-91``: ``yaml.dump_all(..., default_flow_style=False,
explicit_start=True)`` of ``Stream`` objects, keys in PyYAML's sorted order).  Its other format,
HDF5 (``io/hdf5io.py``), needs h5py, which this image does not have.

``PSDFWriter(fname).dump(state)`` takes the particle state where the integrators keep it -- a dict
of device tensors (``Integrator.st.t``, ``BlockHermite`` rows, or a host container with the
reference's attribute names) -- brings the needed columns to the host with ONE copy per column and
streams the text itself: building 10^5-10^6 PyYAML nodes, as ``yaml.dump_all`` does, costs minutes,
while the format is regular enough to be written directly.  Floats are written with ``repr`` so a
snapshot reloads bit for bit.  ``load`` parses a stream back (PyYAML's C-less safe loader with a
constructor for the ``!Particle`` tag) into a dict of numpy arrays -- the test of this module, and the
way back into ``tupan_b200.particles.ParticleSystem``.
"""
import io

import numpy as np

VEC = {"r": ("rx", "ry", "rz"), "v": ("vx", "vy", "vz"), "a": ("ax", "ay", "az")}
SCALARS = (("eps2", "eps2"), ("id", "id"), ("m", "mass"), ("pot", "phi"))     # PSDF key, tupan attribute


def _host(x):
    if hasattr(x, "detach"):                       # torch tensor (any device)
        return x.detach().cpu().numpy()
    return np.asarray(x)


def _columns(state):
    get = (lambda k: state[k]) if isinstance(state, dict) else (lambda k: getattr(state, k))
    has = (lambda k: k in state) if isinstance(state, dict) else (lambda k: hasattr(state, k))
    cols = {}
    for key in ("mass", "eps2", "phi", "time", "tstep", "id") + VEC["r"] + VEC["v"] + VEC["a"]:
        if has(key):
            cols[key] = _host(get(key))
    if "mass" not in cols or not all(k in cols for k in VEC["r"] + VEC["v"]):
        raise ValueError("a snapshot needs mass, rx ry rz, vx vy vz")
    n = len(cols["mass"])
    if "id" not in cols:
        cols["id"] = np.arange(n)
    return n, cols


def _num(x):
    """PyYAML's float representation (repr, with the '.0' / sign conventions it applies)."""
    x = float(x)
    if x != x:
        return ".nan"
    if x in (float("inf"), float("-inf")):
        return ".inf" if x > 0 else "-.inf"
    s = repr(x)
    if "e" in s and "." not in s.split("e")[0]:
        m, e = s.split("e")
        s = m + ".0e" + e                      # PyYAML writes 1e-05 as 1.0e-05 (YAML 1.1 floats need the dot)
    return s


class PSDFWriter(object):
    def __init__(self, fname):
        self.fname = fname

    def dump(self, state, fmode="a", ptype="body", t=None):
        """Append (or write, fmode='w') one snapshot; returns the number of particles written.
        `t` overrides the per-particle time column (shared-step integrators keep one clock)."""
        n, c = _columns(state)
        out = io.StringIO()
        have_a = all(k in c for k in VEC["a"])
        for i in range(n):
            out.write("--- !Particle\n")
            if have_a:
                out.write("a:\n- %s\n- %s\n- %s\n" % tuple(_num(c[k][i]) for k in VEC["a"]))
            if "tstep" in c:
                out.write("dt_next: %s\n" % _num(c["tstep"][i]))
            if "eps2" in c:
                out.write("eps2: %s\n" % _num(c["eps2"][i]))
            out.write("id: %d\n" % int(c["id"][i]))
            out.write("m: %s\n" % _num(c["mass"][i]))
            if "phi" in c:
                out.write("pot: %s\n" % _num(c["phi"][i]))
            out.write("r:\n- %s\n- %s\n- %s\n" % tuple(_num(c[k][i]) for k in VEC["r"]))
            out.write("t: %s\n" % _num(t if t is not None else (c["time"][i] if "time" in c else 0.0)))
            out.write("type: %s\n" % ptype)
            out.write("v:\n- %s\n- %s\n- %s\n" % tuple(_num(c[k][i]) for k in VEC["v"]))
        with open(self.fname, fmode) as f:
            f.write(out.getvalue())
        return n

    def load(self):
        """-> dict of numpy arrays (tupan attribute names) of ALL particles in the stream, in order."""
        import yaml

        class Loader(yaml.SafeLoader):
            pass

        Loader.add_constructor("!Particle", lambda loader, node: loader.construct_mapping(node, deep=True))
        with open(self.fname, "r") as f:
            items = list(yaml.load_all(f, Loader=Loader))
        n = len(items)
        cols = {"id": np.zeros(n, np.int64), "mass": np.zeros(n), "time": np.zeros(n)}
        for key in VEC["r"] + VEC["v"]:
            cols[key] = np.zeros(n)
        opt = {"a": VEC["a"], "eps2": ("eps2",), "pot": ("phi",), "dt_next": ("tstep",)}
        for i, it in enumerate(items):
            cols["id"][i], cols["mass"][i], cols["time"][i] = it["id"], it["m"], it["t"]
            for k, names in (("r", VEC["r"]), ("v", VEC["v"])):
                for name, x in zip(names, it[k]):
                    cols[name][i] = x
            for k, names in opt.items():
                if k in it:
                    vals = it[k] if isinstance(it[k], list) else [it[k]]
                    for name, x in zip(names, vals):
                        cols.setdefault(name, np.zeros(n))[i] = x
        cols["type"] = [it.get("type", "body") for it in items]
        return cols
