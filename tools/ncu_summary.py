#!/usr/bin/env python
"""One line per profiled launch from an ncu report (raw page): the metrics the roofline discussion uses.
usage: tools/ncu_summary.py <report.ncu-rep> [out.csv [steps_captured]]"""
import csv, subprocess, sys, io
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = rows[0]
cols = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum"]
idx = [h.index(c) for c in cols if c in h]
units = rows[1]
out = [[h[i] + (" [" + units[i] + "]" if units[i] else "") for i in idx]]
for r in rows[2:]:
    out.append([r[i].split("(")[0] if h[i] == "Kernel Name" else r[i] for i in idx])
w = csv.writer(open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout)
w.writerows(out)
if len(sys.argv) > 2:
    # tie the capture to the sources it was taken from (bench.py: roofline.dram.same_sources_as_this_library)
    import json, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    json.dump({"csrc_hash": bench.csrc_hash(), "steps_captured": int(sys.argv[3]) if len(sys.argv) > 3 else 1, "report": os.path.basename(sys.argv[1])},
              open(sys.argv[2].replace(".csv", ".meta.json"), "w"))
