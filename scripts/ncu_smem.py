#!/usr/bin/env python
"""Shared-memory wavefronts per SASS instruction of an ncu report (per cell): python scripts/ncu_smem.py rep cells [top]"""
import csv, subprocess, sys
rep, cells = sys.argv[1], float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[1]
isrc, ie, iw, ii, ix = (hdr.index(k) for k in ("Source", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "L1 Wavefronts Shared Excessive"))
tot = 0; items = []; byop = {}
for k, r in enumerate(rows[2:]):
    try: n, w, idl, ex = int(r[ie]), int(r[iw]), int(r[ii]), int(r[ix])
    except (ValueError, IndexError): continue
    if w > 0:
        items.append((w, ex, idl, n, k, r[isrc].strip()[:60]))
        op = r[isrc].split()[1 if r[isrc].strip().startswith("@") else 0].split(".")[0]
        byop[op] = byop.get(op, 0) + w
    tot += w
print("shared-memory wavefronts per cell: %.0f   by opcode: %s" % (tot / cells, "  ".join("%s %.0f" % (k, v / cells) for k, v in sorted(byop.items(), key=lambda kv: -kv[1]))))
items.sort(reverse=True)
for w, ex, idl, n, k, src in items[:top]:
    print("%5d  wavefronts %7.1f (ideal %6.1f, excess %6.1f)  executions %6.1f  %s" % (k, w / cells, idl / cells, ex / cells, n / cells, src))
