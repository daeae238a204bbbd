"""The launch-plan cost model (choose_plan, csrc/pair_engine.cuh) through tupan_cuda_plan_query:
host arithmetic only, runs without a GPU.  The numbers it is checked against are the measured
optima of profiles/r01_plan_probe_*.txt."""
import ctypes

import pytest

from tupan_b200 import backend


def query(lib, kernel, ni, nj, scal=()):
    s = (ctypes.c_double * 8)(*(list(scal) + [0.0] * (8 - len(scal))))
    out = [ctypes.c_int() for _ in range(3)]
    rc = lib.tupan_cuda_plan_query(backend.KERNEL_IDS[kernel], ni, nj, s, *[ctypes.byref(x) for x in out])
    assert rc == 0
    return tuple(x.value for x in out)


@pytest.mark.parametrize("prec", ("float64", "float32"))
def test_every_kernel_gets_a_valid_plan(prec):
    lib = backend.load(prec)
    scal = {"tstep_kernel": (1 / 64,), "pnacc_kernel": (7,) + (1.0,) * 7, "sakura_kernel": (1 / 64, 1),
            "nreg_Xkernel": (0.1,), "nreg_Vkernel": (0.1,)}
    for kernel in backend.KERNEL_IDS:
        if kernel == "kepler_solver_kernel":
            continue
        for ni, nj in ((1, 1), (2, 2), (3, 1000), (1000, 3), (100, 100), (1024, 1024), (5000, 40001),
                       (65536, 65536), (4096, 1 << 20), (1 << 20, 1 << 20), (131072, 1 << 20)):
            split, js, jg = query(lib, kernel, ni, nj, scal.get(kernel, ()))
            assert split in (0, 1, 2) and 0 <= js <= 5 and 1 <= jg <= 64, (kernel, ni, nj, split, js, jg)
            assert jg <= max(1, (nj + 127) // 128), "a chunk holds at least one tile"
            if split != 1:
                assert js == 0
            if split == 2:                       # only kernels with a second group shape may be given it
                assert kernel in ("acc_jerk_kernel", "acc_kernel", "tstep_kernel", "nreg_Xkernel") and prec == "float64"


def test_acc_jerk_fp64_shapes_follow_the_measurements():
    lib = backend.load("float64")
    # tiny systems: lanes of a warp share a particle, one chunk (no finalize launch)
    for n in (2, 64, 512, 1024):
        split, js, jg = query(lib, "acc_jerk_kernel", n, n)
        assert split == 1 and js >= 3, (n, split, js, jg)
    # N = 4096: 16 chunks x 8 i-blocks of 512 on 128 SMs (57 us; the 32-lane split took 190 us) -- the
    # second group shape (2): i-blocks of 768 would leave a third of the last block empty
    split, js, jg = query(lib, "acc_jerk_kernel", 4096, 4096)
    assert split == 2 and 8 <= jg <= 32
    # N = 16384: best measured 32 chunks; 9 chunks (two 15-tile CTAs on most SMs) was 5 % slower
    # (both group shapes are within 1 % of each other here: 22 i-blocks x 13 chunks or 32 x 32)
    split, js, jg = query(lib, "acc_jerk_kernel", 16384, 16384)
    assert split in (0, 2) and jg >= 12
    # large N and the 8-GPU shard of N = 2^20: the 3 x 2 shape (768 particles per CTA), enough chunks to
    # level the SMs (1366 i-blocks x 17 chunks = 156.9 CTAs per SM)
    for ni in (1 << 20, 131072):
        split, js, jg = query(lib, "acc_jerk_kernel", ni, 1 << 20)
        assert split == 0 and jg >= 8
    # rectangular slow<-fast kick of a hierarchical SIA step: few particles against many rows
    split, js, jg = query(lib, "acc_kernel", 16, 65536)
    assert split == 1 and jg > 1
