#!/usr/bin/env bash
# A/B sweep of build variants on the bench workload (device-resident step only): default + names of alt_*.so
set -u
mkdir -p gpurun_out
for v in default "$@" default; do
  if [ "$v" = default ]; then unset JR_B200_LIB; else export JR_B200_LIB=$PWD/jaxrenderer_b200/lib/alt_$v.so; fi
  python bench.py --steps 60 --no-cpu --no-fwd-bwd --no-secondary --e2e eager 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$v ms %.4f median %.4f' % (d['ms_per_step'], d['roofline']['launch_ms_median']))"
done
