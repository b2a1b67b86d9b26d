#!/usr/bin/env bash
# threshold sweep of the tiled raster (variants built with tools/build_variant.sh): cfg 3 / cfg 4 timings only
set -u
for v in default ${EXTRA_VARIANTS:-}; do
  if [ $v = default ]; then unset JR_B200_LIB; else export JR_B200_LIB=$PWD/jaxrenderer_b200/lib/alt_$v.so; fi
  for c in "4 --batch 256" "5 --batch 512"; do
    echo "== $v cfg $c"; timeout 300 python tools/bench_configs.py --cfg $c --steps 5 2>&1 | tail -1 | cut -c1-300
  done
done
echo done
