#!/usr/bin/env bash
# evidence pass after the tiled-raster rework: smoke, tests, both bench arms, ncu --set full of the dominant kernel
# (-> profiles/roofline_traffic.json) and of the tiled raster, launch list of the bench, per-kernel tables of cfg 3 / 4 / 5
set -u
mkdir -p gpurun_out
P=${1:-r2f3}
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${P}_smoke.log; tail -2 gpurun_out/${P}_smoke.log
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${P}_pytest.log; tail -3 gpurun_out/${P}_pytest.log
grep -n "brax frame\|unexplained (" gpurun_out/${P}_pytest.log | head -12
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${P}_bench_reference.json 2> gpurun_out/${P}_bench_reference.err
timeout 900 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; tail -2 gpurun_out/${P}_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vis3 -s 4 -c 1 -o gpurun_out/${P}_vis3 \
  python bench.py --no-cpu --no-fwd-bwd --no-secondary --e2e eager --steps 3 --warmup 3 > gpurun_out/${P}_ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3 -c 60 --csv --log-file gpurun_out/${P}_launches_bench.csv \
  python bench.py --no-cpu --no-fwd-bwd --no-secondary --e2e eager --steps 10 --warmup 3 > gpurun_out/${P}_l.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"
timeout 600 ncu --metrics $M --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/${P}_kernels_cfg5.csv \
  python tools/bench_configs.py --cfg 5 --batch 128 --steps 2 > gpurun_out/${P}_cfg5_ncu.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/${P}_kernels_cfg4.csv \
  python tools/bench_configs.py --cfg 4 --batch 64 --steps 2 > gpurun_out/${P}_cfg4_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_raster_tile -s 6 -c 2 -o gpurun_out/${P}_raster \
  python tools/bench_configs.py --cfg 4 --batch 64 --steps 1 > gpurun_out/${P}_raster_ncu.log 2>&1
python tools/bench_configs.py --cfg 4 --batch 64 --steps 5 > gpurun_out/${P}_cfg4_b64.log 2>&1
python tools/bench_facade.py --batch 4096 --steps 20 > gpurun_out/${P}_facade_4096.json 2>/dev/null
ls -la gpurun_out/${P}_* | wc -l
echo done
