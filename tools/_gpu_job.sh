#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err; echo "rc=$? lines=$(wc -l < gpurun_out/r2f_bench_default.json)"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2f_bench_default.json").read().strip().split("\n")[-1])
print({k:d[k] for k in ("metric","value","unit","n_gpus","steps","warmup","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"].get("traffic"), d["roofline"]["executed"].get("frac_executed"), d["cpu_baseline"]["value"], d["clocks"])
PY
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | head -c 400
