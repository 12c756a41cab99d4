"""Summarise an .ncu-rep (read here on the CPU box): python scripts/ncu_summary.py rep [out.txt]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__cycles_active.avg", "gpc__cycles_elapsed.avg.per_second",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fp64.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"]
out = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in keys:
        if k in d:
            out.append(f"{k:75s} {d[k]:>20s} {units[hdr.index(k)]}")
    out.append("-- warp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active / pcsamp) --")
    st = [(h, d[h]) for h in hdr if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct")]
    st = sorted(st, key=lambda kv: -float(kv[1].replace(",", "") or 0))
    for h, v in st[:10]:
        out.append(f"{h:75s} {v:>20s}")
    fp = [(h, d[h]) for h in hdr if ("fp64" in h or "dmma" in h.lower()) and h not in keys]
    for h, v in fp[:12]:
        out.append(f"{h:75s} {v:>20s}")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(f"# summary of {rep} (ncu --set full --clock-control none)\n" + txt + "\n")
