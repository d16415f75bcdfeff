#!/usr/bin/env python
"""Per-source-line executed warp-instructions of one kernel (ncu source page joined with nvdisasm -g), with the text
of each line.  python tools/sass_lines.py source.csv disasm.sass kernel-substring file.cu nlines [min]"""
import csv, re, sys, collections
src, sass, fn, cu, nlines = sys.argv[1:6]
minv = float(sys.argv[6]) if len(sys.argv) > 6 else 0.1
nlines = float(nlines)
rows = list(csv.reader(open(src))); hdr = rows[1]
ia, ie = hdr.index("Address"), hdr.index("Instructions Executed")
ins = [(int(r[ia], 16), int(r[ie] or 0)) for r in rows[2:] if len(r) > ie and r[ia].startswith("0x")]
base = ins[0][0]
cur = None; inside = False; info = {}
for ln in open(sass):
    if ln.startswith(".text.") and ln.rstrip().endswith(":"):
        inside = fn in ln; continue
    if not inside: continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*);', ln)
    if m: info[int(m.group(1), 16)] = cur
acc = collections.Counter()
for a, e in ins:
    k = info.get(a - base, ("?", 0))
    if k[0] == cu.split("/")[-1]: acc[k[1]] += e
lines = open(cu).read().split("\n")
tot = 0
for l in sorted(acc):
    tot += acc[l]
    if acc[l] / nlines > minv: print("%4d %5.2f  %s" % (l, acc[l] / nlines, lines[l - 1].strip()[:120]))
print("file total %.2f per line" % (tot / nlines))
