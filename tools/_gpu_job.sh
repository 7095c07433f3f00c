#!/bin/bash
# scratch GPU job of the current experiment (edited per call)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
STEPS=20 BENCH_ARGS="--warmup 5" bash tools/run_variants.sh > gpurun_out/r2aa_variants_c3.log 2>&1
cat gpurun_out/r2aa_variants_c3.log
STEPS=6 BENCH_ARGS="--workload acrobot" bash tools/run_variants.sh > gpurun_out/r2aa_variants_acrobot.log 2>&1
cat gpurun_out/r2aa_variants_acrobot.log
