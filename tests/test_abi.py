"""CPU tests: the CUDA libraries load and export every symbol include/libtupan_cuda.h
declares; the Python adapter has the reference's kernel protocol; nothing falls back to CPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from tupan_b200 import backend, extensions, particles

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "libtupan_cuda.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(?:void|int|long long)\s+(\w+)\s*\(", text)))


def ensure_built():
    if not (os.path.exists(backend.lib_path("float64")) and os.path.exists(backend.lib_path("float32"))):
        from tupan_b200 import build
        build.build()


@pytest.mark.parametrize("prec", ("float64", "float32"))
def test_library_exports_every_declared_symbol(prec):
    ensure_built()
    names = declared_functions()
    assert len(names) >= 30
    for ref_name in backend.SIGNATURES:      # the ten entry points of the reference's libtupan.h
        assert ref_name in names
    lib = ctypes.CDLL(backend.lib_path(prec))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert backend.load(prec).tupan_cuda_real_bytes() == (8 if prec == "float64" else 4)


def test_every_new_entry_point_has_a_ctypes_signature():
    # Parts 2, 2b, 3, 3b of the header (everything that is not one of the reference's ten functions)
    # are bound with explicit argument types in backend.PART2: a missing entry would let ctypes guess
    # int-sized arguments for pointers
    names = [n for n in declared_functions() if n not in backend.SIGNATURES]
    missing = [n for n in names if n not in backend.PART2]
    assert not missing, missing
    stale = [n for n in backend.PART2 if n not in names]
    assert not stale, stale


def test_signatures_match_oracle_bindings():
    import oracle
    assert oracle.SIGNATURES == backend.SIGNATURES


def test_adapter_protocol_matches_reference_kernel_objects():
    ensure_built()
    k = backend.CUDAKernel("float64", "acc_jerk_kernel")
    for attr in ("cty", "set_gsize", "set_args", "run", "map_buffers"):
        assert hasattr(k, attr)
    assert k.cty._fields == ("c_int", "c_int_p", "c_uint", "c_uint_p", "c_real", "c_real_p")
    with pytest.raises(TypeError):
        k.cty.c_real_p(np.zeros(4, np.float32))       # wrong dtype must not be passed through
    with pytest.raises(TypeError):
        k.cty.c_real_p(np.zeros(8)[::2])              # nor a strided view


def test_extension_marshalling_order():
    ensure_built()
    ps = particles.ParticleSystem(5)
    ext = extensions.AccJerk("CUDA", "float64")
    ext.set_args(ps, ps[:3])
    assert ext._inargs[0] == 5 and ext._inargs[9] == 3
    assert ext._inargs[1] is ps.mass and ext._inargs[5] is ps.eps2 and ext._inargs[8] is ps.vz
    assert [a is getattr(ps, n) for a, n in zip(ext._outargs, ("ax", "ay", "az", "jx", "jy", "jz"))] == [True] * 6
    with pytest.raises(ValueError):
        extensions.Acc("C", "float64")


def test_no_cpu_fallback_without_gpu():
    ensure_built()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ps = particles.ParticleSystem(4)
    ps.mass[...] = 1.0
    ps.rx[...] = np.arange(4)
    with pytest.raises(backend.TupanCudaError):
        ps.set_acc(ps)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tupan_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "libtupan_oracle" not in text and "libtupan_ref" not in text, f
