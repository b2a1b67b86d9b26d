#!/usr/bin/env bash
# quick check: tests + facade timing + per-kernel table of the facade (+ optional cfg timings)
set -u
mkdir -p gpurun_out
P=${1:-r2q}
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${P}_pytest.log; tail -4 gpurun_out/${P}_pytest.log
grep -n "FAILED\|Error" gpurun_out/${P}_pytest.log | head -10
python tools/bench_facade.py --batch 4096 --steps 20 > gpurun_out/${P}_facade_4096.json 2>/dev/null; cut -c1-300 gpurun_out/${P}_facade_4096.json
for c in 3 4 5; do python tools/bench_configs.py --cfg $c --steps 5 > gpurun_out/${P}_cfg$c.log 2>&1; tail -1 gpurun_out/${P}_cfg$c.log | cut -c1-200; done
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"
timeout 600 ncu --metrics $M --clock-control none -k regex:^k_ -s 24 -c 24 --csv --log-file gpurun_out/${P}_kernels_facade.csv \
  python tools/bench_facade.py --batch 4096 --steps 3 > gpurun_out/${P}_facade_ncu.log 2>&1
python tools/ncu_kernel_table.py gpurun_out/${P}_kernels_facade.csv 2>/dev/null | cut -c1-160
echo done
