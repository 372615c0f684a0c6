#!/usr/bin/env python
"""Key figures of an ncu report (first kernel): python scripts/ncu_summary.py rep.ncu-rep [cells]"""
import csv, subprocess, sys
rep = sys.argv[1]; cells = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
def g(k):
    return float(d[k][0].replace(",", "")) if k in d and d[k][0] not in ("", "n/a") else float("nan")
print("kernel:", d.get("Kernel Name", ("?",))[0][:100])
t = g("gpu__time_duration.sum"); tu = d["gpu__time_duration.sum"][1]
print("duration %.4f %s   grid %s block %s regs %s  dyn smem %s" % (t, tu, d["launch__grid_size"][0], d["launch__block_size"][0], d["launch__registers_per_thread"][0], d.get("launch__shared_mem_per_block_dynamic", ("?",))[0]))
inst = g("smsp__inst_executed.sum")
print("warp-instructions %.4g%s" % (inst, "  = %.0f per cell" % (inst / cells) if cells else ""))
print("issue active %.1f%%   warps active %.1f%% of peak  achieved warps/scheduler %.2f" % (g("smsp__issue_active.avg.pct_of_peak_sustained_active"), g("sm__warps_active.avg.pct_of_peak_sustained_active"), g("smsp__warps_active.avg.per_cycle_active")))
rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
print("dram read %.1f %s write %.1f %s  (%.1f%% of peak)%s" % (rd, d["dram__bytes_read.sum"][1], wr, d["dram__bytes_write.sum"][1], g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ""))
print("l1 hit %.1f%%  l2 hit %.1f%%" % (g("l1tex__t_sector_hit_rate.pct"), g("lts__t_sector_hit_rate.pct")))
print("smem wavefronts %.4g (conflicts %.4g: ld %.3g st %.3g atom %.3g)" % (g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"), g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"), g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"), g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum")))
st = [(k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), g(k)) for k in d if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
st.sort(key=lambda kv: -kv[1])
print("stalls per issue: " + ", ".join("%s %.2f" % kv for kv in st[:9]))
