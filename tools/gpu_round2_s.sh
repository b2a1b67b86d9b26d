#!/usr/bin/env bash
# tests + B=512 / B=4096 bench (graph launch), optional A/B through JR_NO_CLUSTER
set -u
mkdir -p gpurun_out
P=${1:-r2s2}
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${P}_pytest.log; tail -4 gpurun_out/${P}_pytest.log
grep -n "FAILED\|Error" gpurun_out/${P}_pytest.log | head -10
for nb in 512 256 128 1024; do
for nc in 0 1; do
  if [ $nc = 1 ]; then export JR_NO_CLUSTER=1; else unset JR_NO_CLUSTER; fi
  timeout 300 python bench.py --steps 50 --batch $nb --launch graph --no-cpu --no-fwd-bwd --no-secondary --e2e eager > gpurun_out/${P}_b${nb}_nc$nc.json 2> gpurun_out/${P}_b${nb}_nc$nc.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${P}_b${nb}_nc$nc.json').read().strip().splitlines()[-1])
print('B=$nb no_cluster=$nc ms %.4f median %.4f images/s %.0f' % (d['ms_per_step'], d['roofline']['launch_ms_median'], d['value']))
PY
done; done
unset JR_NO_CLUSTER
echo done
