#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2i_smoke.log; tail -3 gpurun_out/r2i_smoke.log
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -5 gpurun_out/r2i_pytest.log
grep -n "FAILED\|Error" gpurun_out/r2i_pytest.log | head -20
timeout 900 python bench.py --steps 30 --no-secondary --no-fwd-bwd --no-cpu > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -3 gpurun_out/r2i_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'])
print('e2e',d['e2e'])
print('fp32',{k:v for k,v in d['roofline']['fp32'].items() if k!='counters'})
print('fwd_bwd',d.get('fwd_bwd'))
print('secondary',json.dumps(d.get('secondary'),indent=1))
print('cpu',d.get('cpu_baseline'))
PY
