#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2c_bench_${tag}_8gpu.json 2> gpurun_out/r2c_bench_${tag}_8gpu.err
  python - <<PY
import json
t=[l for l in open("gpurun_out/r2c_bench_${tag}_8gpu.json").read().split("\n") if l.startswith("{")]
d=json.loads(t[-1]); print("$tag 8 GPUs", "%.4g"%d["value"], "ms %.3f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "frac", d["roofline"]["frac"], "agents/GPU", d["config"].get("agents_per_gpu"))
PY
}
run c3
run c3_1M --agents 131072
run c5 --workload c5
run c4 --workload c4
