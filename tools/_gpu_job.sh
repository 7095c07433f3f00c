#!/bin/bash
# scratch GPU job of the current experiment (edited per call)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r2ae_pytest.log 2>&1
tail -3 gpurun_out/r2ae_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2ae_bench_c3.json 2> gpurun_out/r2ae_bench_c3.err; tail -c 600 gpurun_out/r2ae_bench_c3.json
for w in acrobot cartpole c4 c5; do python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2ae_bench_$w.json 2> gpurun_out/r2ae_bench_$w.err; python -c "
import json,sys; d=json.load(open('gpurun_out/r2ae_bench_$w.json')); print('$w', '%.4g'%d['value'], 'ms %.3f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'])"; done
