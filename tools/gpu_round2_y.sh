#!/usr/bin/env bash
# visible-triangle lists built by the tiled raster's resolve: tests + A/B against the k_mark_visible launch
set -u
mkdir -p gpurun_out
P=${1:-r2y}
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${P}_pytest.log; tail -3 gpurun_out/${P}_pytest.log
for v in fused separate; do
  if [ $v = separate ]; then export JR_NO_FUSED_MARK=1; else unset JR_NO_FUSED_MARK; fi
  for c in "4 --batch 256" "5 --batch 512"; do
    echo "== $v cfg $c"; timeout 300 python tools/bench_configs.py --cfg $c --steps 5 2>&1 | tail -1 | cut -c1-260
  done
done
unset JR_NO_FUSED_MARK
echo done
