#!/usr/bin/env bash
# short evidence pass: tests, bench arms, ncu --set full of the dominant kernel (-> roofline_traffic.json), cfg 3 / 4 timings
set -u
mkdir -p gpurun_out
P=${1:-r2f5}
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${P}_pytest.log; tail -3 gpurun_out/${P}_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vis3 -s 4 -c 1 -o gpurun_out/${P}_vis3 \
  python bench.py --no-cpu --no-fwd-bwd --no-secondary --e2e eager --steps 3 --warmup 3 > gpurun_out/${P}_ncu_full.log 2>&1
timeout 900 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; tail -2 gpurun_out/${P}_bench.err
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"
timeout 600 ncu --metrics $M --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/${P}_kernels_cfg4.csv \
  python tools/bench_configs.py --cfg 4 --batch 64 --steps 2 > gpurun_out/${P}_cfg4_ncu.log 2>&1
echo done
