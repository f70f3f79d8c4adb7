#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.txt]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum ", "lts__t_bytes.sum.per_second",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ", "smsp__average_warp_latency_per_inst_issued",
        "smsp__average_warps_issue_stalled", "sass__inst_executed_local", "sm__inst_executed_pipe_uniform", "sm__inst_executed_pipe_xu", "smsp__inst_executed_pipe_",
        "l1tex__t_sector_hit_rate", "sm__inst_executed_pipe_adu", "sm__inst_executed_pipe_cbu", "sm__cycles_active.avg ",
        "sm__pipe_fp64_cycles_active.avg.pct", "sm__inst_executed_pipe_fp64.avg.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct", "l1tex__data_pipe_lsu_wavefronts.sum ",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum ", "l1tex__m_xbar2l1tex_read_bytes.sum ", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"]
out = []
for r in rows[2:]:
    out.append(f"== {r[hdr.index('Kernel Name')]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
    for h, u, v in zip(hdr, units, r):
        if any(h.startswith(w.strip()) if w.endswith(" ") else w in h for w in WANT) and "max_rate" not in h:
            out.append(f"{h} [{u}] = {v}")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
