#!/usr/bin/env python
"""Summarise an .ncu-rep: per kernel launch key metrics (raw page) -> text table.  python scripts/ncu_summary.py rep [out]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = [
    ("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
    ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%"), ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_noinst"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__inst_executed_op_local_ld.sum", "local_ld"), ("smsp__inst_executed_op_local_st.sum", "local_st"),
    ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "local_ld_sectors"), ("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "local_st_sectors"),
]
idx = [(hdr.index(k), n) for k, n in want if k in hdr]
out = []
for r in rows[2:]:
    parts = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = v.replace("ncb::", "").replace("(ncb::NarrowArgs)", "")[:44]
        else:
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            v = f"{n}={v}{units[i] if n in ('time','dram_rd','dram_wr') else ''}"
        parts.append(v)
    out.append("  ".join(parts))
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
