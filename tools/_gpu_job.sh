#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2an_pytest.log 2>&1
tail -15 gpurun_out/r2an_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c3', '%.4g'%d['value'], 'ms %.3f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'])"
