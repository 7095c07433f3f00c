#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/r2aw_pytest.log 2>&1
tail -6 gpurun_out/r2aw_pytest.log
