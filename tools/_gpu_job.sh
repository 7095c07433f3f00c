#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/r2c_multi_gpu_tests.log 2>&1
cat gpurun_out/r2c_multi_gpu_tests.log
