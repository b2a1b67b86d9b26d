#!/usr/bin/env bash
# ncu --set full + source of the secondary hot kernels: shading (facade), tiled raster (cfg 4), backward pixel pass (cfg 5)
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade_rec -s 3 -c 1 -o gpurun_out/r2o_shade_rec \
  python tools/bench_facade.py --batch 4096 --steps 2 > gpurun_out/r2o_shade.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_raster_tile -s 4 -c 2 -o gpurun_out/r2o_raster_tile \
  python tools/bench_configs.py --cfg 4 --batch 64 --steps 2 > gpurun_out/r2o_raster.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bwd_global -s 3 -c 1 -o gpurun_out/r2o_bwd_global \
  python tools/bench_configs.py --cfg 5 --batch 128 --steps 2 > gpurun_out/r2o_bwd.log 2>&1
ls -la gpurun_out/r2o_*
echo done
