#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log; tail -3 gpurun_out/r2l_pytest.log
for v in default $@; do
  if [ "$v" = default ]; then unset JR_B200_LIB; else export JR_B200_LIB=$PWD/jaxrenderer_b200/lib/alt_$v.so; fi
  timeout 300 python bench.py --no-cpu --no-fwd-bwd --no-secondary --steps 50 > gpurun_out/r2l_bench_$v.json 2> gpurun_out/r2l_bench_$v.err
  python -c "
import json
for line in open('gpurun_out/r2l_bench_$v.json'):
    if line.startswith('{'):
        d=json.loads(line);print('$v', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), round(d['e2e']['value']))"
done
unset JR_B200_LIB
echo skipncu

echo done
