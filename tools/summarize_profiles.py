#!/usr/bin/env python3
"""Turn the raw ncu outputs a gpurun call leaves in gpurun_out/ (r1_launches.csv, r1_k_run_steps.ncu-rep, r1_bench.json) into
the tracked summaries under profiles/ (launch shares, details page, selected raw metrics, DRAM traffic per launch)."""
import collections, csv, json, os, re, shutil, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
G, P = "gpurun_out", "profiles"
shutil.copy(f"{G}/{tag}_launches.csv", f"{P}/{tag}_launches.csv")
rows = list(csv.reader(open(f"{G}/{tag}_launches.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(h)]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    name = re.sub(r"\(.*", "", r[ik]); v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
with open(f"{P}/{tag}_launch_shares.csv", "w") as f:
    f.write("kernel,launches,total_us,share\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{v[0]},{v[1]:.1f},{v[1] / tot:.4f}\n")
open(f"{P}/{tag}_k_run_steps_details.txt", "w").write(subprocess.run(["ncu", "-i", f"{G}/{tag}_k_run_steps.ncu-rep", "--page", "details"], capture_output=True, text=True).stdout)
raw = subprocess.run(["ncu", "-i", f"{G}/{tag}_k_run_steps.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); h = rows[0]; units = dict(zip(h, rows[1]))
keep = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "sm__cycles_active.avg",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
out, tr = [], []
for r in rows[2:]:
    d = dict(zip(h, r)); o = {k: (d.get(k), units.get(k)) for k in keep}
    o["stalls_per_issue"] = sorted(((float(x), n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                                    for n, x in d.items() if "issue_stalled" in n and n.endswith("per_issue_active.ratio")), reverse=True)[:8]
    out.append(o)
    b = lambda k: float(o[k][0]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[o[k][1]]
    tr.append(b("dram__bytes_read.sum") + b("dram__bytes_write.sum"))
json.dump(out, open(f"{P}/{tag}_k_run_steps_metrics.json", "w"), indent=1)
json.dump({"k_run_steps": {"dram_bytes_per_launch": sum(tr) / len(tr), "launches_profiled": len(tr),
                           "source": f"profiles/{tag}_k_run_steps_metrics.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)"}},
          open(f"{P}/traffic.json", "w"), indent=1)
if os.path.exists(f"{G}/{tag}_bench.json"):
    shutil.copy(f"{G}/{tag}_bench.json", f"{P}/{tag}_bench.json")
print(open(f"{P}/{tag}_launch_shares.csv").read())
for o in out:
    print(o["gpu__time_duration.sum"], "issue", o["smsp__issue_active.avg.per_cycle_active"][0], "fp64 pipe", o["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0],
          "warps", o["smsp__warps_active.avg.per_cycle_active"][0], "regs", o["launch__registers_per_thread"][0], o["stalls_per_issue"][:3])
