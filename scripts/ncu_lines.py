#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line:  python scripts/ncu_lines.py gpurun_out/prof.ncu-rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; items = []; tot_i = 0; tot_s = 0
cur = "?"
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 5: continue
    if r[2] != "-": continue          # per-SASS rows; keep the per-source-line aggregate rows only
    try:
        n = int(r[hdr.index("Instructions Executed")]); s = int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    items.append((n, s, cur[:14] + ":" + r[0], r[1].strip()[:100])); tot_i += n; tot_s += s
items.sort(reverse=True)
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
for n, s, l, src in items[:top]:
    print("%12d %5.1f%%  stall-samples %5.1f%%  %-20s %s" % (n, 100.0 * n / max(tot_i, 1), 100.0 * s / max(tot_s, 1), l, src))
