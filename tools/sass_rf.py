#!/usr/bin/env python
"""Static register-file-read count of the FP64 instructions in a kernel's hot loop.

    cuobjdump -sass -fun <mangled> file.o | python tools/sass_rf.py [--loop N]

Model (fitted to tools/microbench2.cu on a B200): an FP64 instruction occupies the pipe for 2
clocks; a DFMA whose three source operands are three distinct registers, none served by the
operand-reuse cache (the previous instruction of the warp read the same register in the same
slot with .reuse), needs a third clock to collect its operands.  The script finds the largest
backward-branch loop, and prints the instruction mix and the modelled clocks per iteration.
"""
import re
import sys
from collections import Counter

INS = re.compile(r'^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);')


def parse(lines):
    out = []
    for l in lines:
        m = INS.match(l)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip()))
    return out


def operands(text):
    # "DFMA R80, R86.reuse, R62, R18" -> op, [dst, srcs...]
    if text.startswith('@'):
        text = text.split(None, 1)[1]
    op, _, rest = text.partition(' ')
    args = [a.strip() for a in rest.split(',')] if rest else []
    return op, args


def reg_of(a):
    m = re.match(r'^[-|~]?\|?(R\d+)', a)
    return m.group(1) if m else None


def stats(ins, mufu_per_pair=1):
    """dict(dp, three, other, clocks, pairs) of the largest loop; pairs = MUFU count (one rsqrt per pair)."""
    r = _loop_stats(ins)
    if r is None:
        return None
    r["pairs"] = max(r["mix"].get("MUFU", 0) / float(mufu_per_pair), 1.0)
    return r


def analyse(ins):
    r = _loop_stats(ins)
    if r is None:
        print("no loop found")
        return
    mix, dp, other, three, clocks, n = r["mix"], r["dp"], r["other"], r["three"], r["clocks"], r["n"]
    print("loop: %d instructions, %d FP64 (%s), %d other" % (n, dp, dict((k, mix[k]) for k in ('DFMA', 'DADD', 'DMUL')), other))
    print("other:", {k: v for k, v in mix.items() if k not in ('DFMA', 'DADD', 'DMUL')})
    print("DFMAs with 3 uncached register operands: %d" % three)
    print("modelled FP64-pipe clocks per iteration: %.0f (2 x %d = %d if none needed a third clock)" % (clocks, dp, 2 * dp))
    return dp, three, other


def _loop_stats(ins):
    # largest loop = backward BRA with the most instructions
    loops = []
    for k, (addr, text) in enumerate(ins):
        op, args = operands(text)
        if op.startswith('BRA') and args:
            try:
                tgt = int(args[-1], 16)
            except ValueError:
                continue
            if tgt <= addr:
                j = next(i for i, (a, _) in enumerate(ins) if a >= tgt)
                loops.append((j, k))
    # the hot loop: the loop with the most FP64 instructions among those that contain a MUFU (one
    # rsqrt seed per pair) while none of the loops nested in them does (so a loop split into several
    # basic blocks by small inner loops is taken whole)
    def n_of(j, k, names):
        return sum(1 for _, t in ins[j:k + 1] if operands(t)[0].split('.')[0] in names)
    best = None
    for (j, k) in loops:
        if n_of(j, k, ('MUFU',)) == 0:
            continue
        inner = [(j2, k2) for (j2, k2) in loops if (j2, k2) != (j, k) and j <= j2 and k2 <= k]
        if any(n_of(j2, k2, ('MUFU',)) > 0 for (j2, k2) in inner):
            continue
        ndp = n_of(j, k, ('DFMA', 'DADD', 'DMUL'))
        if best is None or ndp > best[2]:
            best = (j, k, ndp)
    if best is None:
        return None
    body = ins[best[0]:best[1] + 1]
    # rarely executed blocks: the range a per-thread predicated forward branch jumps over when that
    # range holds no MUFU (the mask fix-up of a grouped kernel: `@!P0 BRA` over the block that re-forms
    # r2).  The uniform fences (`BRA.U !UP0`) of the basic blocks are never taken and are not skipped.
    skip = set()
    end_addr = body[-1][0]
    for addr, text in body:
        m = re.match(r'^@!?P\d\s+BRA\s+(?:.*\s)?0x([0-9a-f]+)$', text)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if addr < tgt <= end_addr:
            rng = [(a, t) for a, t in body if addr < a < tgt]
            if rng and not any(operands(t)[0].startswith('MUFU') for _, t in rng):
                skip.update(a for a, _ in rng)
    rare = len(skip)
    body = [(a, t) for a, t in body if a not in skip]
    mix = Counter()
    dp = 0
    clocks = 0.0
    three = 0
    prev_reuse = {}
    for addr, text in body:
        op, args = operands(text)
        base = op.split('.')[0]
        mix[base] += 1
        srcs = args[1:]
        cur_reuse = {}
        if base in ('DFMA', 'DADD', 'DMUL'):
            dp += 1
            regs = set()
            for slot, a in enumerate(srcs):
                r = reg_of(a)
                if r is None or r == 'RZ':
                    continue
                if prev_reuse.get(slot) == r:
                    continue                     # served by the reuse cache
                regs.add(r)
            c = 2.0
            if base == 'DFMA' and len(regs) >= 3:
                c = 3.0
                three += 1
            clocks += c
        for slot, a in enumerate(srcs):
            if '.reuse' in a:
                cur_reuse[slot] = reg_of(a)
        prev_reuse = cur_reuse
    other = sum(v for k, v in mix.items() if k not in ('DFMA', 'DADD', 'DMUL'))
    return {"mix": mix, "dp": dp, "other": other, "three": three, "clocks": clocks, "n": len(body), "rare": rare}


if __name__ == "__main__":
    analyse(parse(sys.stdin.readlines()))
