#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_bench_c3_8gpu.json 2> gpurun_out/r2e_bench_c3_8gpu.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2e_bench_c3_8gpu.json").read().strip().split("\n")[-1]); print('c3 8 GPUs', '%.4g'%d['value'], 'ms %.3f'%d['ms_per_step'], 'e2e ms %.3f'%d['e2e']['ms_per_step'], d.get('ms_per_step_by_rank'))
PY
echo "stdout lines $(wc -l < gpurun_out/r2e_bench_c3_8gpu.json)"
