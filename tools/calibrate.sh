#!/bin/bash
# one ncu pass per fused kernel: FP64 thread instructions + DRAM bytes of ONE 125-step launch, beside the launch's device-side work counters
cd "$(dirname "$0")/.."
M=smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum
for w in c3 acrobot cartpole; do
  skip=40; [ $w != c3 ] && skip=16
  ncu --metrics $M --clock-control none -k regex:k_run_steps --launch-skip $skip --launch-count 1 --csv --log-file gpurun_out/cal_$w.csv \
      python tools/calibrate.py run $w > gpurun_out/cal_$w.log 2>&1
  grep CALIB gpurun_out/cal_$w.log
done
