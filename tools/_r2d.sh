python -m pytest tests -m gpu -x -q -k "shared or retiled" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench_c3.json 2> gpurun_out/r2d_bench_c3.err
for w in c4 c5 c1; do
  python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_$w.json 2> gpurun_out/r2d_bench_$w.err
done
tail -3 gpurun_out/r2d_pytest.log
