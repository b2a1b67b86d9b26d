#!/usr/bin/env bash
# tests + full bench line (secondary lines with per-stage rooflines from the library's kernel timing)
set -u
mkdir -p gpurun_out
P=${1:-r2t}
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${P}_pytest.log; tail -4 gpurun_out/${P}_pytest.log
grep -n "FAILED\|Error" gpurun_out/${P}_pytest.log | head -10
timeout 900 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${P}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${P}_bench.json').read().strip().splitlines()[-1])
print('value %.0f ms %.4f e2e %.0f frac %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))
for k,v in (d.get('secondary') or {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk != 'stages'})
    for st in v.get('stages', []):
        print('   ', st['stage'], '%.4f ms' % st['ms'], st['kernels_ms'], 'frac %.3f' % st['roofline']['frac'] if 'roofline' in st else '')
for k,v in (d.get('fwd_bwd') or {}).items():
    print(k, v['value'], v.get('kernels_ms'))
PY
echo done
