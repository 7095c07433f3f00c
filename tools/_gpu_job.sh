#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
STEPS=6 BENCH_ARGS="--workload acrobot" bash tools/run_variants.sh 2>&1 | tee gpurun_out/r2ax_wupd.log
STEPS=6 BENCH_ARGS="--workload c5" bash tools/run_variants.sh 2>&1 | tee -a gpurun_out/r2ax_wupd.log
