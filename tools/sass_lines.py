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
        lines.append((cur, m.group(2).strip()))
rows = list(csv.reader(open(srccsv)))
hdr = rows[1]; ia = hdr.index("Instructions Executed"); ist = hdr.index("Warp Stall Sampling (All Samples)")
cnt = [(int(r[ia]), int(r[ist] or 0)) for r in rows[2:] if len(r) > ia and r[ia].isdigit()]
print("sass instrs:", len(lines), "ncu rows:", len(cnt))
agg = collections.Counter(); st = collections.Counter()
for (loc, txt), (n, s) in zip(lines, cnt):
    agg[loc] += n; st[loc] += s
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
