python -m pytest tests -m gpu -x -q -k "policy or table or exploration or bench_workload or team or merge or golden or full_size" > gpurun_out/r2g_pytest.log 2>&1; tail -3 gpurun_out/r2g_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline | grep '^{' > gpurun_out/r2g_c3.json
python bench.py --workload acrobot --steps 6 --warmup 3 --no-cpu-baseline | grep '^{' > gpurun_out/r2g_acrobot.json
python bench.py --workload c5 --steps 6 --warmup 3 --no-cpu-baseline | grep '^{' > gpurun_out/r2g_c5.json
