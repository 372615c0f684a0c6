#!/usr/bin/env python
"""Per-SASS-instruction view of an ncu report: opcode histogram (per cell) and the listing with executed counts.
   python scripts/ncu_sass.py rep.ncu-rep cells [out.txt]"""
import csv, collections, subprocess, sys
rep, cells = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, ie, isamp, ithr = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Predicated-On Threads Executed")
ops = collections.Counter(); tot = 0; data = []
for r in rows[2:]:
    try: n = int(r[ie])
    except (ValueError, IndexError): continue
    toks = r[isrc].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    ops[op.split(".")[0]] += n; tot += n; data.append((r[ia][-5:], r[isrc].strip(), n, r[isamp], r[ithr]))
print("warp-instructions per cell: %.0f" % (tot / cells))
print("  ".join("%s %.0f" % (k, v / cells) for k, v in ops.most_common(28)))
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write("\n".join("%s %9.1f %6s %5s  %s" % (a, n / cells, s, t, src) for a, src, n, s, t in data))
