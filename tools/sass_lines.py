#!/usr/bin/env python
"""Attribute ncu per-SASS-instruction execution counts to CUDA source lines.
usage: sass_lines.py <lib.so> <kernel mangled substring> <ncu source csv> <units(frames)>"""
import csv, collections, re, subprocess, sys, os, tempfile
so, ksub, srccsv, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "ofdm_io" not in f][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
# locate kernel section
lines = []; inside = False; cur = None
for l in sass:
    if l.startswith("//---------------------"):
        inside = (".text." in l) and (ksub in l); continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        lines.append((cur, m.group(2).strip(), int(m.group(1), 16)))
rows = list(csv.reader(open(srccsv)))
hdr = rows[1]; ia = hdr.index("Instructions Executed"); ist = hdr.index("Warp Stall Sampling (All Samples)")
body = [r for r in rows[2:] if len(r) > ia and r[ia].isdigit()]
base = int(body[0][0], 16)
cnt = {int(r[0], 16) - base: (int(r[ia]), int(r[ist] or 0), r[hdr.index("Source")]) for r in body}
print("sass instrs:", len(lines), "ncu rows:", len(cnt))
agg = collections.Counter(); st = collections.Counter()
# the profiled build may differ by a few instructions from the library on disk: align the two opcode streams
import difflib
def opc(t):
    w = t.split()
    return (w[1] if w[0].startswith("@") and len(w) > 1 else w[0])
A = [opc(t) for _, t, _ in lines]
B = [opc(r[hdr.index("Source")]) for r in body]
sm = difflib.SequenceMatcher(None, A, B, autojunk=False)
matched = 0
for blk in sm.get_matching_blocks():
    for i in range(blk.size):
        loc = lines[blk.a + i][0]; r = body[blk.b + i]
        agg[loc] += int(r[ia]); st[loc] += int(r[ist] or 0); matched += 1
print("aligned instructions:", matched, "of", len(A), "/", len(B))
srcs = {}
tot = sum(agg.values())
print("total per unit: %.1f" % (tot / units))
for loc, n in sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0))):
    if n / units < 1.0: continue
    f, ln = loc if loc else ("?", 0)
    if f not in srcs:
        p = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[f][ln - 1].strip()[:100] if 0 < ln <= len(srcs[f]) else ""
    print("%7.1f st=%5d %s:%d  %s" % (n / units, st[loc], f, ln, text))
