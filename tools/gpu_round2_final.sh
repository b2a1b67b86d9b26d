#!/usr/bin/env bash
# final evidence pass of round 2: tests, smoke, both bench arms, ncu --set full of the dominant kernel and of the shading
# kernel, launch list of the bench, per-kernel tables of the facade / cfg 3 / 4 / 5
set -u
mkdir -p gpurun_out
P=${1:-r2z}
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${P}_smoke.log; tail -2 gpurun_out/${P}_smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${P}_pytest.log; tail -3 gpurun_out/${P}_pytest.log
timeout 900 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; tail -2 gpurun_out/${P}_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${P}_bench_reference.json 2> gpurun_out/${P}_bench_reference.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vis3 -s 4 -c 1 -o gpurun_out/${P}_vis3 \
  python bench.py --no-cpu --no-fwd-bwd --no-secondary --e2e eager --steps 3 --warmup 3 > gpurun_out/${P}_ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3 -c 60 --csv --log-file gpurun_out/${P}_launches_bench.csv \
  python bench.py --no-cpu --no-fwd-bwd --no-secondary --e2e eager --steps 10 --warmup 3 > gpurun_out/${P}_l.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"
timeout 600 ncu --metrics $M --clock-control none -k regex:^k_ -s 21 -c 21 --csv --log-file gpurun_out/${P}_kernels_facade.csv \
  python tools/bench_facade.py --batch 4096 --steps 3 > gpurun_out/${P}_facade_ncu.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/${P}_kernels_cfg5.csv \
  python tools/bench_configs.py --cfg 5 --batch 128 --steps 2 > gpurun_out/${P}_cfg5_ncu.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/${P}_kernels_cfg4.csv \
  python tools/bench_configs.py --cfg 4 --batch 64 --steps 2 > gpurun_out/${P}_cfg4_ncu.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -s 9 -c 18 --csv --log-file gpurun_out/${P}_kernels_cfg3.csv \
  python tools/bench_configs.py --cfg 3 --steps 2 > gpurun_out/${P}_cfg3_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade_rec -s 3 -c 1 -o gpurun_out/${P}_shade_rec \
  python tools/bench_facade.py --batch 4096 --steps 2 > gpurun_out/${P}_shade_ncu.log 2>&1
python tools/bench_facade.py --batch 4096 --steps 20 --merged > gpurun_out/${P}_facade_4096.json 2>/dev/null
python tools/bench_facade.py --batch 1024 --steps 20 --graph > gpurun_out/${P}_facade_1024.json 2>/dev/null
for c in 3 4 5; do python tools/bench_configs.py --cfg $c --steps 5 > gpurun_out/${P}_cfg$c.log 2>&1; done
python tools/bench_configs.py --cfg 4 --batch 256 --steps 3 > gpurun_out/${P}_cfg4_b256.log 2>&1
python tools/bench_configs.py --cfg 5 --batch 512 --steps 3 > gpurun_out/${P}_cfg5_b512.log 2>&1
ls -la gpurun_out/${P}_* | wc -l
echo done
