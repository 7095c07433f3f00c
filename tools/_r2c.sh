python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_c3.json 2> gpurun_out/r2c_bench_c3.err
for w in acrobot cartpole c4 c5 c1 c2; do
  python bench.py --workload $w --steps 6 --warmup 3 > gpurun_out/r2c_bench_$w.json 2> gpurun_out/r2c_bench_$w.err
done
tail -3 gpurun_out/r2c_pytest.log
