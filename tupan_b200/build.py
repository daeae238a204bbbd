"""Build the two CUDA libraries in-tree with nvcc for sm_100a.

    python -m tupan_b200.build            # both precisions, parallel over translation units

Outputs ``tupan_b200/lib/libtupan_cuda_fp{32,64}.so`` (git-ignored, shipped to the GPU box
by gpurun).  Objects are cached under ``tupan_b200/lib/obj`` and rebuilt when a source or
header is newer; each object's ``ptxas -v`` report is cached next to it and
``tupan_b200/lib/ptxas.log`` (git-ignored; a copy per round is kept under ``profiles/``) is
assembled from all of them on every build, cached or not.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
UNITS = ["abi", "context", "peer", "k_newton", "k_accjerk", "k_snap", "k_nreg", "k_pn", "k_sakura", "k_update"]
# k_update evaluates the O(N) integrator updates in the reference's numpy operation order, one
# rounding per operation: no FMA contraction in that unit.
# k_accjerk: the register-usage level that suits the grouped acc_jerk kernel (see the unit's header).
# k_snap: same flag, +0.8 % for the grouped snap_crackle kernel.
_RUL10 = ["-Xptxas", "-regUsageLevel", "-Xptxas", "10"]
UNIT_FLAGS = {"k_update": ["-fmad=false"], "k_accjerk": _RUL10, "k_snap": _RUL10}
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _newest_header():
    t = 0.0
    for root in (SRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".h")):
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def _compile(args):
    unit, tag, hdr_time, verbose = args
    src = os.path.join(SRC, unit + ".cu")
    obj = os.path.join(LIBDIR, "obj", "%s_%s.o" % (unit, tag))
    log = obj[:-2] + ".ptxas.log"
    if os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_time):
        # the ptxas report (registers, spills) of a cached object is kept next to it
        return unit, tag, "cached", open(log).read() if os.path.exists(log) else ""
    cmd = ([NVCC] + FLAGS + UNIT_FLAGS.get(unit, []) + os.environ.get("TUPAN_NVCC_EXTRA", "").split()
           + (["-DTUPAN_FP64"] if tag == "fp64" else [])
           + ["-c", src, "-o", obj])
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s (%s):\n%s\n%s" % (unit, tag, p.stdout, p.stderr))
    with open(log, "w") as f:
        f.write(p.stderr)
    return unit, tag, "built", p.stderr


def build(verbose=False, jobs=None):
    os.makedirs(os.path.join(LIBDIR, "obj"), exist_ok=True)
    hdr_time = _newest_header()
    work = [(u, tag, hdr_time, verbose) for tag in ("fp64", "fp32") for u in UNITS]
    logs = []
    with ThreadPoolExecutor(jobs or os.cpu_count() or 4) as ex:
        for unit, tag, what, log in ex.map(_compile, work):
            logs.append((unit, tag, log))
            if verbose:
                print("[%s %s] %s" % (unit, tag, what))
    with open(os.path.join(LIBDIR, "ptxas.log"), "w") as f:
        for unit, tag, log in logs:
            if log:
                f.write("==== %s %s ====\n%s\n" % (unit, tag, log))
    for tag in ("fp64", "fp32"):
        out = os.path.join(LIBDIR, "libtupan_cuda_%s.so" % tag)
        objs = [os.path.join(LIBDIR, "obj", "%s_%s.o" % (u, tag)) for u in UNITS]
        if os.path.exists(out) and os.path.getmtime(out) > max(os.path.getmtime(o) for o in objs):
            continue
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed (%s):\n%s\n%s" % (tag, p.stdout, p.stderr))
    return [os.path.join(LIBDIR, "libtupan_cuda_%s.so" % t) for t in ("fp64", "fp32")]


if __name__ == "__main__":
    for path in build(verbose=True):
        print(path)
