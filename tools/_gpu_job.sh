#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "odd_capacity or full_rule_bases or overflow" 2>&1 | tail -15 ) > gpurun_out/r2ao_pytest.log 2>&1
tail -15 gpurun_out/r2ao_pytest.log
