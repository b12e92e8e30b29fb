#!/usr/bin/env python
"""Static SASS instruction mix per kernel of libofdm_b200.so (cuobjdump -sass): the instruction-level evidence kept
under profiles/ (TMA bulk copies UBLKCP, mbarrier SYNCS, packed fp32 FADD2/FMUL2/FFMA2, conversions F2F, FP64 ops).

    python tools/sass_count.py [lib.so] [substring ...]  > profiles/rN_sass_counts.txt
"""
import collections
import re
import subprocess
import sys

WATCH = ["UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "F2F", "DMUL", "DADD", "DFMA", "MUFU",
         "LDS", "STS", "LDG", "STG", "SHFL", "MOV", "IMAD", "LOP3", "PRMT", "BAR", "REDUX", "UTMALDG", "UTMASTG", "LDL", "STL"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1].endswith(".so") else "ieee-802.11-ofdm-qpsk-simulator_b200/libofdm_b200.so"
    pats = [a for a in sys.argv[1:] if not a.endswith(".so")]
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
    kernels, cur = [], None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = [names[len(kernels)], collections.Counter()]
            kernels.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            if op != "NOP":
                cur[1][op] += 1
                cur[1]["total"] += 1
    print("%-72s %6s  %s" % ("kernel (static SASS, whole function incl. rare paths)", "total", "  ".join("%s" % w for w in WATCH)))
    for name, c in sorted(kernels, key=lambda k: k[0]):
        short = name.replace("ofdm::", "").replace("void ", "").replace("(int)", "").replace("(bool)", "")
        short = re.sub(r">\(.*$", ">", short) if ">(" in short else re.sub(r"\(.*$", "", short)
        if pats and not any(p in short for p in pats):
            continue
        print("%-72s %6d  %s" % (short[:72], c["total"], "  ".join("%*d" % (len(w), c[w]) for w in WATCH)))


if __name__ == "__main__":
    main()
