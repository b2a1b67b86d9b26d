#!/usr/bin/env bash
# tests + bench of the default library (+ optional variants: names of alt_*.so)
set -u
mkdir -p gpurun_out
P=${1:-r2r}; shift || true
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${P}_pytest.log; tail -4 gpurun_out/${P}_pytest.log
grep -n "FAILED\|Error" gpurun_out/${P}_pytest.log | head -10
for v in default "$@"; do
  if [ "$v" = default ]; then unset JR_B200_LIB; else export JR_B200_LIB=$PWD/jaxrenderer_b200/lib/alt_$v.so; fi
  timeout 300 python bench.py --steps 50 --no-cpu --no-fwd-bwd --no-secondary --e2e eager > gpurun_out/${P}_bench_$v.json 2> gpurun_out/${P}_bench_$v.err
  tail -2 gpurun_out/${P}_bench_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${P}_bench_$v.json').read().strip().splitlines()[-1])
print('$v value %.0f ms %.4f median %.4f frac %.4f' % (d['value'], d['ms_per_step'], d['roofline']['launch_ms_median'], d['roofline']['frac']))
PY
  timeout 300 python bench.py --steps 50 --batch 512 --no-cpu --no-fwd-bwd --no-secondary --e2e eager > gpurun_out/${P}_bench512_$v.json 2> gpurun_out/${P}_bench512_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${P}_bench512_$v.json').read().strip().splitlines()[-1])
print('$v B=512 value %.0f ms %.4f median %.4f' % (d['value'], d['ms_per_step'], d['roofline']['launch_ms_median']))
PY
done
echo done
