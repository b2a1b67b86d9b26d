#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -4 gpurun_out/r2g_pytest.log
for v in default c5 g4 g16 default; do
  if [ "$v" = default ]; then unset JR_B200_LIB; else export JR_B200_LIB=$PWD/jaxrenderer_b200/lib/alt_$v.so; fi
  timeout 300 python bench.py --no-cpu --no-fwd-bwd --steps 50 > gpurun_out/r2g_bench_$v.json 2> gpurun_out/r2g_bench_$v.err
  python -c "import json;d=json.load(open('gpurun_out/r2g_bench_$v.json'));print('$v', d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
done
unset JR_B200_LIB
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vis3 -s 4 -c 1 -o gpurun_out/r2g_vis3 \
  python bench.py --no-cpu --no-fwd-bwd --steps 3 --warmup 3 > gpurun_out/r2g_ncu.log 2>&1
echo done
