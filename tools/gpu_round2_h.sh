#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -5 gpurun_out/r2h_pytest.log
grep -n "FAILED\|Error\|cfg4\|cfg5\|grad position\|visibility counters" gpurun_out/r2h_pytest.log | head -40
timeout 300 python bench.py --no-cpu --no-fwd-bwd --steps 50 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
python -c "import json;d=json.load(open('gpurun_out/r2h_bench.json'));print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
