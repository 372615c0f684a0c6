#!/usr/bin/env python
"""Sum an ncu report's executed warp-instructions and stall samples over source-line ranges:
   python scripts/ncu_phase_sum.py rep.ncu-rep file.cuh name:lo-hi name:lo-hi ...   (per-cell figures with --cells N)"""
import csv, subprocess, sys
rep, fname = sys.argv[1], sys.argv[2]
ranges = []; cells = 1
for a in sys.argv[3:]:
    if a.startswith("--cells="): cells = int(a.split("=")[1]); continue
    nm, r = a.split(":"); lo, hi = r.split("-"); ranges.append((nm, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur_file = None; acc = {nm: [0, 0] for nm, _, _ in ranges}; other = [0, 0]; tot = [0, 0]
hdr_other = {}
for r in rows:
    if r and r[0] == "File Path": cur_file = r[1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None:
        continue
    if len(r) < len(hdr) - 5: continue
    if r[2] != "-": continue
    try:
        n = int(r[hdr.index("Instructions Executed")]); s = int(r[hdr.index("# Samples")]); ln = int(r[0])
    except ValueError:
        continue
    tot[0] += n; tot[1] += s
    if cur_file is None or not cur_file.endswith(fname):
        k = (cur_file or "?").split("/")[-1]
        hdr_other.setdefault(k, [0, 0]); hdr_other[k][0] += n; hdr_other[k][1] += s
        continue
    hit = False
    for nm, lo, hi in ranges:
        if lo <= ln <= hi:
            acc[nm][0] += n; acc[nm][1] += s; hit = True; break
    if not hit: other[0] += n; other[1] += s
print("total: %.0f warp-instr/cell, %d samples" % (tot[0] / cells, tot[1]))
for nm, lo, hi in ranges:
    print("%-14s L%4d-%-4d %9.0f instr/cell %5.1f%%   stall samples %5.1f%%" % (nm, lo, hi, acc[nm][0] / cells, 100.0 * acc[nm][0] / tot[0], 100.0 * acc[nm][1] / tot[1]))
print("%-14s %20.0f instr/cell %5.1f%%   stall samples %5.1f%%" % ("other", other[0] / cells, 100.0 * other[0] / tot[0], 100.0 * other[1] / tot[1]))
for k, v in sorted(hdr_other.items(), key=lambda kv: -kv[1][0]):
    print("%-34s %9.0f instr/cell %5.1f%%   stall samples %5.1f%%" % (k, v[0] / cells, 100.0 * v[0] / tot[0], 100.0 * v[1] / tot[1]))
