#!/usr/bin/env bash
# compute-sanitizer over the reworked tiled raster (queue of warp-cooperative triangles, span raster)
set -u
mkdir -p gpurun_out
P=${1:-r2w}
timeout 700 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_named_configs.py tests/test_gpu_edge_cases.py -m gpu -q -x \
  -k "long_needles or odd_canvas or straddling or many_large or kernel_timing or binned or 480" > gpurun_out/${P}_memcheck.log 2>&1; tail -4 gpurun_out/${P}_memcheck.log
timeout 700 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_edge_cases.py -m gpu -q -x \
  -k "(long_needles and (case0 or case2)) or straddling or many_large" > gpurun_out/${P}_racecheck.log 2>&1; tail -4 gpurun_out/${P}_racecheck.log
echo done
