#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
STEPS=20 BENCH_ARGS="--warmup 5 --policy 3 --groups 1" bash tools/run_variants.sh 2>&1 | tee gpurun_out/r2av_sliced.log
STEPS=6 BENCH_ARGS="--workload c5 --policy 3 --groups 1" bash tools/run_variants.sh 2>&1 | tee -a gpurun_out/r2av_sliced.log
