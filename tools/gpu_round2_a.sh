#!/usr/bin/env bash
# first GPU pass of round 2: parity of the new single-tile kernel, bench A/B against k_vis2, ncu
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 300 python bench.py --no-cpu --no-fwd-bwd --steps 50 > gpurun_out/r2a_bench_vis3.json 2> gpurun_out/r2a_bench_vis3.err
JR_VIS2=1 timeout 300 python bench.py --no-cpu --no-fwd-bwd --steps 50 > gpurun_out/r2a_bench_vis2.json 2> gpurun_out/r2a_bench_vis2.err
cat gpurun_out/r2a_bench_vis3.json gpurun_out/r2a_bench_vis2.json | cut -c1-600
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vis3 -s 4 -c 1 -o gpurun_out/r2a_vis3 \
  python bench.py --no-cpu --no-fwd-bwd --steps 3 --warmup 3 > gpurun_out/r2a_ncu.log 2>&1
echo done
