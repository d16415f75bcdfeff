#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS dump with `nvdisasm -g -c` line info: executed warp-instructions
per (file, line) and per file.  Tuning aid.
    python tools/sass_by_line.py source.csv disasm.sass <mangled-name-substring> [topN]"""
import csv, re, sys, collections

def main():
    src, sass, fn = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    ia, ie, iss = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Source")
    ismp = hdr.index("# Samples")
    ins = [(int(r[ia], 16), int(r[ie] or 0), r[iss].strip(), int(r[ismp] or 0)) for r in rows[2:] if len(r) > ie and r[ia].startswith("0x")]
    base = ins[0][0]
    # line info of the function
    cur = None; inside = False; info = {}
    for ln in open(sass):
        if ln.startswith(".text.") and ln.rstrip().endswith(":"):
            inside = fn in ln
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
        m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*);', ln)
        if m:
            info[int(m.group(1), 16)] = cur
    by_line = collections.Counter(); by_file = collections.Counter(); smp_line = collections.Counter()
    total = 0; tsmp = 0
    for a, e, s, sm in ins:
        k = info.get(a - base, ("?", 0))
        by_line[k] += e; by_file[k[0]] += e; total += e; smp_line[k] += sm; tsmp += sm
    print("total warp-instructions executed: %d, samples %d" % (total, tsmp))
    for f, e in by_file.most_common():
        print("  %-28s %12d  %5.1f%%" % (f, e, 100.0 * e / total))
    print("top lines:")
    for (f, l), e in by_line.most_common(top):
        print("  %-24s:%-5d %12d  %5.1f%%   samples %5.1f%%" % (f, l, e, 100.0 * e / total, 100.0 * smp_line[(f, l)] / max(tsmp, 1)))

main()
