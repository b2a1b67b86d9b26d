#!/usr/bin/env bash
# tiled raster A/B: span raster of the sparse boxes (default build) vs the hierarchical block raster (alt_nospan.so)
set -u
mkdir -p gpurun_out
P=${1:-r2u}
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${P}_pytest.log; tail -4 gpurun_out/${P}_pytest.log
grep -n "FAILED\|Error" gpurun_out/${P}_pytest.log | head -10
for v in default ${EXTRA_VARIANTS:-}; do
  if [ $v = default ]; then unset JR_B200_LIB; else export JR_B200_LIB=$PWD/jaxrenderer_b200/lib/alt_$v.so; fi
  for c in "4 --batch 256" "5 --batch 512"; do
    echo "== $v cfg $c"
    timeout 300 python tools/bench_configs.py --cfg $c --steps 5 2>&1 | tail -3
  done
done
unset JR_B200_LIB
echo done
