"""Static properties of the shipped fp64 kernels, read from the SASS of the built library (no GPU):
the instruction counts DESIGN.md section 4 argues with.  tools/sass_rf.py finds the hot loop of a kernel
and counts FP64-pipe instructions, DFMAs that have to collect three distinct registers (none served by
the operand-reuse cache: one extra pipe clock each, tools/microbench2.cu) and everything else, per pair.
A compiler flag, a header edit or a ptxas change that silently breaks the grouping shows up here."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sass_rf  # noqa: E402

LIB = os.path.join(ROOT, "tupan_b200", "lib", "libtupan_cuda_fp64.so")

# kernel (demangled prefix) -> (MUFU per pair, max FP64 per pair, max uncollected DFMAs per pair, max other per pair)
EXPECT = {
    "pair_kernel_grouped<tupan::AccJerkOp<double>, 256, 3, 2, 27, 128, 4, false>": (1, 31.0, 2.1, 6.5),
    "pair_kernel_grouped<tupan::AccJerkOp<double>, 256, 2, 4, 27, 128, 4, false>": (1, 31.0, 3.6, 6.8),
    "pair_kernel_grouped<tupan::AccOp<double>, 256, 6, 2, 8, 128, 4, false>": (1, 18.0, 1.1, 5.0),
    "pair_kernel_grouped<tupan::PhiOp<double>, 256, 3, 2, 8, 128, 4, false>": (1, 13.0, 0.8, 6.0),
    "pair_kernel_grouped<tupan::TstepOp<double>, 256, 3, 2, 8, 128, 4, false>": (2, 36.0, 3.0, 13.0),
    "pair_kernel_grouped<tupan::NregXOp<double>, 256, 4, 2, 8, 128, 4, false>": (1, 28.0, 2.3, 6.0),
    "pair_kernel_grouped<tupan::SnapCrackleOp<double>, 256, 1, 4, 1, 128, 4, false>": (1, 78.0, 10.0, 15.0),
}


@pytest.fixture(scope="module")
def functions():
    if not (shutil.which("cuobjdump") and shutil.which("c++filt")):
        pytest.skip("needs cuobjdump and c++filt")
    if not os.path.exists(LIB):
        pytest.skip("library not built")
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, timeout=600).stdout
    out = {}
    chunks = re.split(r'\n\s*Function : ', txt)[1:]
    names = [c.split('\n')[0].strip() for c in chunks]
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    for name, chunk in zip(dem, chunks):
        out[re.sub(r'^void ', '', re.sub(r'\(.*', '', name)).replace("tupan::pair_kernel", "pair_kernel")] = chunk
    return out


@pytest.mark.parametrize("kernel", sorted(EXPECT))
def test_hot_loop_instruction_mix(functions, kernel):
    mufu, dp_max, three_max, other_max = EXPECT[kernel]
    assert kernel in functions, "kernel not in the library: %s" % kernel
    r = sass_rf.stats(sass_rf.parse(functions[kernel].split('\n')), mufu_per_pair=mufu)
    assert r is not None, "no hot loop found"
    n = r["pairs"]
    assert r["dp"] / n <= dp_max + 1e-9, ("FP64 instructions per pair", r["dp"] / n)
    assert r["three"] / n <= three_max, ("uncollected three-register DFMAs per pair", r["three"] / n)
    assert r["other"] / n <= other_max, ("non-FP64 instructions per pair", r["other"] / n)


def test_headline_kernel_has_no_spills():
    log = os.path.join(ROOT, "tupan_b200", "lib", "ptxas.log")
    if not os.path.exists(log):
        pytest.skip("no ptxas log (library built elsewhere)")
    txt = open(log).read()
    blocks = re.findall(r"Function properties for (\S*pair_kernel_grouped\S*)\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores", txt)
    assert blocks, "no grouped kernel in the ptxas log"
    for name, stack, spill in blocks:
        assert int(spill) == 0 and int(stack) == 0, (name, stack, spill)
