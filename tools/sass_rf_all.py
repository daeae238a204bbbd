#!/usr/bin/env python
"""Run tools/sass_rf.py's model over every kernel of a binary whose name matches a pattern.
    python tools/sass_rf_all.py tools/bin/kernel_lab2 lab_kernel
Prints one line per kernel: FP64 per innermost-loop pair body, 3-register DFMAs, other instructions.
"""
import re
import subprocess
import sys
sys.path.insert(0, __import__("os").path.dirname(__file__))
import sass_rf  # noqa: E402


def main():
    binary, pat = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    txt = subprocess.run(["cuobjdump", "-sass", binary], capture_output=True, text=True).stdout
    for chunk in re.split(r'\n\s*Function : ', txt)[1:]:
        name = chunk.split('\n')[0].strip()
        if pat and pat not in name:
            continue
        ins = sass_rf.parse(chunk.split('\n'))
        r = sass_rf.stats(ins)
        if r is None:
            continue
        tag = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        tag = re.sub(r'\(.*', '', tag)
        print("%-70s pairs/loop %5.1f  FP64/pair %5.2f  3-reg DFMA/pair %5.2f  other/pair %5.2f  model clk/pair %6.2f%s"
              % (tag[-70:], r["pairs"], r["dp"] / r["pairs"], r["three"] / r["pairs"], r["other"] / r["pairs"],
                 r["clocks"] / r["pairs"],
                 "  (+%d instructions in rarely taken blocks, not counted)" % r["rare"] if r.get("rare") else ""))


if __name__ == "__main__":
    main()
