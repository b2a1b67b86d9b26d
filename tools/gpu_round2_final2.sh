#!/usr/bin/env bash
# last pass: tests, smoke, both bench arms, ncu --set full of the dominant kernel (-> roofline_traffic.json), launch list
set -u
mkdir -p gpurun_out
P=${1:-r2zz}
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${P}_smoke.log; tail -2 gpurun_out/${P}_smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${P}_pytest.log; tail -3 gpurun_out/${P}_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vis3 -s 4 -c 1 -o gpurun_out/${P}_vis3 \
  python bench.py --no-cpu --no-fwd-bwd --no-secondary --e2e eager --steps 3 --warmup 3 > gpurun_out/${P}_ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3 -c 60 --csv --log-file gpurun_out/${P}_launches_bench.csv \
  python bench.py --no-cpu --no-fwd-bwd --no-secondary --e2e eager --steps 10 --warmup 3 > gpurun_out/${P}_l.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_forward.py -m gpu -q -x -k "clustered or visible_triangle or brax_fixture or instanced or depth_brax_like" > gpurun_out/${P}_memcheck.log 2>&1; tail -5 gpurun_out/${P}_memcheck.log
echo done
