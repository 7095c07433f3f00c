#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/r2at_pytest.log 2>&1
tail -2 gpurun_out/r2at_pytest.log
STEPS=20 BENCH_ARGS="--warmup 5" bash tools/run_variants.sh 2>&1 | tee gpurun_out/r2at_qpre.log
STEPS=6 BENCH_ARGS="--workload acrobot" bash tools/run_variants.sh 2>&1 | tee -a gpurun_out/r2at_qpre.log
STEPS=6 BENCH_ARGS="--workload c5" bash tools/run_variants.sh 2>&1 | tee -a gpurun_out/r2at_qpre.log
