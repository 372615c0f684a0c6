#!/usr/bin/env python
"""Group ncu per-line instruction counts of fused_kernel.cuh into kernel phases (markers in the source)."""
import csv, subprocess, sys, re
rep = sys.argv[1]; cells = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; per = {}
for r in rows:
    if r and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 8 or r[2] != "-": continue
    try: n = int(r[hdr.index("Instructions Executed")]); s = int(r[hdr.index("# Samples")])
    except ValueError: continue
    per[(cur_file, int(r[0]))] = (n, s, r[1])
SRC = sys.argv[3] if len(sys.argv) > 3 else "fused_kernel.cuh"
src = open("pyscenic_b200/csrc/" + SRC).read().splitlines()
# phase boundaries by searching marker text
marks = [("scan_helper", "block_excl_scan_n("), ("scatter_packed", "void scatter_packed("), ("write_aucs", "void write_aucs_packed("),
         ("general_cell", "void general_cell("), ("slot_dispatch(inlined phase bodies)", "struct SlotIdx"), ("row_stream", "struct RowView"), ("kernel_prologue", "aucell_fused_kernel(const FusedArgs"), ("LOAD", "---- LOAD ----"),
         ("stats", "cell statistics"), ("S1_hist1", "// S1:"), ("S2_mixed_scan", "// S2:"), ("S3_clear", "// S3:"), ("S4_hist2", "// S4:"),
         ("S5_scan2", "// S5:"), ("T1_classify", "// T1:"), ("T2_fill", "// T2:"), ("T3_scan", "// T3:"), ("T4_rank", "// T4:"), ("T4_small", "// small bins: one thread per entry"), ("S8_scatter", "// S8:"), ("emit", "auto emit = "), ("S6_place", "// S6:"), ("S7_rank", "// S7:"), ("ZEROS", "ZEROS: fill ranks"), ("tail", "if (a.stats != nullptr && tid == 0) {"),
         ("from_rankings", "aucell_from_rankings_kernel(")]
if SRC == "plane_kernel.cuh":
    marks = [("prologue", "aucell_plane_kernel(const PlaneArgs"), ("LOAD", "---- LOAD ----"), ("stats_reduce", "c_exp = warp_sum(c_exp);"),
             ("S1_hist1", "// S1:"), ("S2_mixed", "// S2:"), ("S2_scan", "constexpr int per = (kNb1"), ("T1_classify", "// T1:"),
             ("T2_fill", "// T2:"), ("T3_scan", "// T3:"), ("T4_rank", "// T4:"), ("T4_small", "if (tid < nsmall) {"),
             ("SCATTER_expand", "// SCATTER:"), ("todo", "if (todo) {"), ("ZEROS", "---- ZEROS"), ("OUT", "---- OUT ----")]
bounds = []
for name, text in marks:
    for i, l in enumerate(src):
        if text in l: bounds.append((i + 1, name)); break
bounds.sort()
def phase(line):
    p = "header"
    for b, n in bounds:
        if line >= b: p = n
    return p
tot = sum(v[0] for v in per.values()); agg = {}
for (f, ln), (n, s, t) in per.items():
    k = phase(ln) if f == SRC else f
    a = agg.setdefault(k, [0, 0]); a[0] += n; a[1] += s
tots = sum(v[1] for v in agg.values())
print("total warp-instr %.3g (%.0f per cell)" % (tot, tot / cells))
for k, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-18s %6.1f%% instr  %8.0f /cell   %5.1f%% stall samples" % (k, 100.0 * n / tot, n / cells, 100.0 * s / max(tots, 1)))
