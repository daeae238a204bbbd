"""PSDF snapshot writer (tupan_b200/psdf.py; reference format: tupan/io/psdfio.py:25-91).

* the text written equals, byte for byte, what the reference's own recipe produces for the same
  particles -- ``yaml.dump_all(objects, default_flow_style=False, explicit_start=True)`` of
  ``!Particle`` objects with the reference's attribute set (restated here with PyYAML, so the test
  runs without the reference);
* a snapshot reloads bit for bit;
* where the reference is installed (baseline/_ref or /root/reference) its own ``Stream.to_dumper`` is
  run on a reference particle system and compared as well."""
import os
import sys

import numpy as np
import pytest
import yaml

from tupan_b200 import ics
from tupan_b200.psdf import PSDFWriter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Particle(yaml.YAMLObject):
    yaml_tag = "!Particle"

    @classmethod
    def to_yaml(cls, dumper, data):               # psdfio.py:69-91
        return dumper.represent_mapping(data.yaml_tag, data.__dict__)


def reference_text(ps, with_a, t):
    objs = []
    for i in range(ps.n):
        o = Particle()
        o.id, o.m, o.t = int(ps.id[i]), float(ps.mass[i]), float(t)
        o.r = [float(ps.rx[i]), float(ps.ry[i]), float(ps.rz[i])]
        o.v = [float(ps.vx[i]), float(ps.vy[i]), float(ps.vz[i])]
        if with_a:
            o.a = [float(ps.ax[i]), float(ps.ay[i]), float(ps.az[i])]
        o.type, o.eps2 = "body", float(ps.eps2[i])
        objs.append(o)
    return yaml.dump_all(objs, default_flow_style=False, explicit_start=True)


def test_text_equals_pyyaml_dump_of_the_reference_recipe(tmp_path):
    ps = ics.make_plummer(37, seed=2)
    ps.id = np.arange(ps.n)
    rng = np.random.default_rng(0)
    for k in ("ax", "ay", "az"):
        setattr(ps, k, rng.standard_normal(ps.n) * 10.0 ** rng.integers(-12, 12, ps.n))
    ps.rx[0], ps.ry[0], ps.vz[0] = 1e-5, -3.0, 1e22          # exponent / integer-valued float spellings
    f = str(tmp_path / "snap.psdf")
    st = {k: getattr(ps, k) for k in ("id", "mass", "eps2", "rx", "ry", "rz", "vx", "vy", "vz", "ax", "ay", "az")}
    assert PSDFWriter(f).dump(st, fmode="w", t=0.25) == ps.n
    assert open(f).read() == reference_text(ps, True, 0.25)


def test_round_trip_is_bit_exact_and_appends(tmp_path):
    ps = ics.make_plummer(64, seed=3)
    f = str(tmp_path / "snap.psdf")
    st = {k: getattr(ps, k) for k in ("mass", "eps2", "rx", "ry", "rz", "vx", "vy", "vz")}
    st["time"] = np.full(ps.n, 0.125)
    st["tstep"] = np.full(ps.n, 2.0 ** -7)
    w = PSDFWriter(f)
    w.dump(st, fmode="w")
    w.dump(st)                                               # append, as simulation.py dumps worldlines
    back = w.load()
    assert len(back["mass"]) == 2 * ps.n
    for k in ("mass", "eps2", "rx", "ry", "rz", "vx", "vy", "vz", "time", "tstep"):
        assert np.array_equal(back[k][:ps.n], st[k]) and np.array_equal(back[k][ps.n:], st[k]), k
    assert np.array_equal(back["id"][:ps.n], np.arange(ps.n)) and set(back["type"]) == {"body"}


def test_against_the_reference_stream_class(tmp_path):
    ref = next((p for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference")
                if os.path.isdir(os.path.join(p, "tupan", "io"))), None)
    if ref is None:
        pytest.skip("no reference")
    import subprocess
    script = r'''
import sys, os, tempfile
os.environ["HOME"] = tempfile.mkdtemp()
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np, yaml
from tupan.ics.plummer import make_plummer
# tupan/io/__init__.py imports h5py (absent here): load psdfio.py as tupan.io.psdfio without it
import types, importlib, tupan
pkg = types.ModuleType("tupan.io"); pkg.__path__ = [os.path.join(os.path.dirname(tupan.__file__), "io")]
sys.modules["tupan.io"] = pkg
Stream = importlib.import_module("tupan.io.psdfio").Stream
from tupan_b200.psdf import PSDFWriter
ps = make_plummer(16, 4.0 / 16, ("equalmass",), seed=1)
ps.set_phi(ps); ps.set_acc(ps)
# the reference's per-particle view (obj.pos / obj.vel) predates its SoA containers: feed its
# dumper the same values through the attribute names it reads
class P(object): pass
objs = []
for i in range(ps.n):
    o = Stream()
    o.id, o.m, o.t_curr = int(ps.id[i]), float(ps.mass[i]), float(ps.time[i])
    o.r = [float(ps.rx[i]), float(ps.ry[i]), float(ps.rz[i])]
    o.v = [float(ps.vx[i]), float(ps.vy[i]), float(ps.vz[i])]
    o.a = [float(ps.ax[i]), float(ps.ay[i]), float(ps.az[i])]
    o.type, o.eps2, o.pot = "body", float(ps.eps2[i]), float(ps.phi[i])
    objs.append(o)
want = yaml.dump_all(objs, default_flow_style=False, explicit_start=True)     # Stream.to_yaml, psdfio.py:69-91
f = os.path.join(tempfile.mkdtemp(), "s.psdf")
st = {k: getattr(ps, k) for k in ("id", "mass", "eps2", "phi", "time", "rx", "ry", "rz", "vx", "vy", "vz", "ax", "ay", "az")}
PSDFWriter(f).dump(st, fmode="w")
assert open(f).read() == want, (open(f).read()[:400], want[:400])
print("PSDF-REF-OK")
''' % (ROOT, ref)
    p = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=600)
    assert "PSDF-REF-OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
