#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of a built library (run here, no GPU needed):

    python tools/sass_hist.py neural_spectral_codec_b200/libnsc_b200.so > profiles/sass_<tag>.txt

Prints, for every kernel, the instruction count and the counts of the opcodes that identify how
it moves data and synchronises (UBLKCP = cp.async.bulk / TMA, SYNCS = mbarrier, LDGSTS = cp.async,
ATOMS/RED = shared atomics, FFMA2/FMUL2 = packed FP32, MUFU, BAR, LDG/STG/LDS/STS)."""
import collections
import re
import subprocess
import sys

WATCH = ["UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "LDGDEPBAR", "ATOMS", "RED", "ATOMG", "FFMA2", "FMUL2", "FADD2",
         "FFMA", "MUFU", "BAR", "LDG", "STG", "LDS", "STS", "DFMA", "DADD", "DMUL", "SHFL", "REDUX", "MULTIMEM"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True, check=True).stdout
    name, hist, total = None, collections.Counter(), 0
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()

    def flush():
        if name is None:
            return
        short = re.sub(r"\(anonymous namespace\)::|nsc::", "", demangle(name))
        short = re.sub(r"\(.*", "", short)
        parts = "  ".join(f"{k} {hist[k]}" for k in WATCH if hist[k])
        print(f"{short}\n    {total} instructions:  {parts}")

    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            flush()
            name, hist, total = m.group(1), collections.Counter(), 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and name:
            total += 1
            hist[m.group(1)] += 1
    flush()


if __name__ == "__main__":
    main()
