#!/usr/bin/env bash
# profiling pass: ncu --set full of the dominant kernel, launch list of the bench, per-kernel metrics of the facade / cfg 3 / 4 / 5
set -u
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vis3 -s 4 -c 1 -o gpurun_out/r2m_vis3 \
  python bench.py --no-cpu --no-fwd-bwd --no-secondary --steps 3 --warmup 3 > gpurun_out/r2m_ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3 -c 60 --csv --log-file gpurun_out/r2m_launches_bench.csv \
  python bench.py --no-cpu --no-fwd-bwd --no-secondary --steps 10 --warmup 3 > gpurun_out/r2m_l.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/r2m_kernels_facade.csv \
  python tools/bench_facade.py --batch 4096 --steps 3 > gpurun_out/r2m_facade.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/r2m_kernels_cfg5.csv \
  python tools/bench_configs.py --cfg 5 --batch 128 --steps 2 > gpurun_out/r2m_cfg5.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/r2m_kernels_cfg4.csv \
  python tools/bench_configs.py --cfg 4 --batch 64 --steps 2 > gpurun_out/r2m_cfg4.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -s 12 -c 24 --csv --log-file gpurun_out/r2m_kernels_cfg3.csv \
  python tools/bench_configs.py --cfg 3 --steps 2 > gpurun_out/r2m_cfg3.log 2>&1
python tools/bench_facade.py --batch 4096 --steps 20 > gpurun_out/r2m_facade_timing.json 2>/dev/null; cat gpurun_out/r2m_facade_timing.json | cut -c1-400
python tools/bench_facade.py --batch 1024 --steps 20 > gpurun_out/r2m_facade_timing_1024.json 2>/dev/null; cat gpurun_out/r2m_facade_timing_1024.json | cut -c1-400
tail -2 gpurun_out/r2m_cfg5.log gpurun_out/r2m_cfg4.log gpurun_out/r2m_cfg3.log | cut -c1-300
echo done
