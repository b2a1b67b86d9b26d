#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2n_launches_bench.csv \
  python bench.py --no-cpu --no-fwd-bwd --no-secondary --steps 10 --warmup 3 > gpurun_out/r2n_l.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:^k_ -s 24 -c 24 --csv --log-file gpurun_out/r2n_kernels_facade.csv \
  python tools/bench_facade.py --batch 4096 --steps 3 > gpurun_out/r2n_facade.log 2>&1
echo done
