python -m pytest tests/test_multi_gpu.py -x -q > gpurun_out/r2f_multi.log 2>&1; tail -5 gpurun_out/r2f_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
$TR 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | grep '^{' > gpurun_out/r2f_2gpu_a.json
NCCL_MAX_CTAS=1 $TR 29512 bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | grep '^{' > gpurun_out/r2f_2gpu_b.json
NCCL_MAX_CTAS=1 $TR 29513 bench.py --gpus 2 --steps 20 --warmup 5 --nccl-high-priority 0 2>/dev/null | grep '^{' > gpurun_out/r2f_2gpu_c.json
$TR 29514 bench.py --gpus 2 --steps 20 --warmup 5 --nccl-high-priority 0 2>/dev/null | grep '^{' > gpurun_out/r2f_2gpu_d.json
python bench.py --steps 20 --warmup 5 --no-cpu-baseline | grep '^{' > gpurun_out/r2f_1gpu.json
python bench.py --workload acrobot --steps 6 --warmup 3 --no-cpu-baseline | grep '^{' > gpurun_out/r2f_acrobot.json
python -m pytest tests -m gpu -x -q -k "policy or table or exploration or bench_workload or team or merge" > gpurun_out/r2f_pytest.log 2>&1; tail -3 gpurun_out/r2f_pytest.log
