#!/usr/bin/env python3
"""tools/summarize_ncu.py <gpurun_out/x.ncu-rep> <profiles/name> -- the tracked summary of one `ncu --set full` capture:
<name>_details.txt (details page) and <name>_metrics.json (the raw metrics that matter + stall reasons per issue + DRAM bytes)."""
import csv, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
open(out + "_details.txt", "w").write(subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); h = rows[0]; units = dict(zip(h, rows[1]))
keep = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "sm__cycles_active.avg",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
res = []
for r in rows[2:]:
    d = dict(zip(h, r)); o = {k: (d.get(k), units.get(k)) for k in keep if k in d}
    o["stalls_per_issue"] = sorted(((float(x), n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                                    for n, x in d.items() if "issue_stalled" in n and n.endswith("per_issue_active.ratio") and x), reverse=True)[:8]
    res.append(o)
json.dump(res, open(out + "_metrics.json", "w"), indent=1)
for o in res:
    print({k: v[0] for k, v in o.items() if k != "stalls_per_issue"}); print(o["stalls_per_issue"])
