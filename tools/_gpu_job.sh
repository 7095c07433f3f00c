#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench_c3_2gpu.json 2> gpurun_out/r2d_bench_c3_2gpu.err
echo "stdout lines: $(wc -l < gpurun_out/r2d_bench_c3_2gpu.json)"; head -c 150 gpurun_out/r2d_bench_c3_2gpu.json; echo; grep -c "NCCL version" gpurun_out/r2d_bench_c3_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2d_ref_2gpu.json 2> gpurun_out/r2d_ref_2gpu.err; echo "reference arm stdout lines: $(wc -l < gpurun_out/r2d_ref_2gpu.json)"; head -c 200 gpurun_out/r2d_ref_2gpu.json
