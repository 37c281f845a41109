#!/usr/bin/env python
"""Per-launch balance figures from an ncu report: duration, achieved warps, and SM active cycles avg / max (how much of a
launch is tail).  usage: tools/ncu_balance.py <report.ncu-rep> [out.csv]"""
import csv, io, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = rows[0]
cols = ["Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "sm__cycles_active.avg", "sm__cycles_active.max",
        "sm__cycles_elapsed.max", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__waves_per_multiprocessor", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__inst_executed.sum"]
idx = [h.index(c) for c in cols if c in h]
out = [[h[i] + (" [" + rows[1][i] + "]" if rows[1][i] else "") for i in idx]]
for r in rows[2:]:
    out.append([r[i].split("(")[0] if h[i] == "Kernel Name" else r[i] for i in idx])
csv.writer(open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout).writerows(out)
