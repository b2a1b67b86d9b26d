#!/usr/bin/env bash
# multi-GPU pass (final): NCCL gradient equivalence, host-link ceiling, weak / strong scaling of the bench
set -u
N=${1:-2}
P=${2:-r2y}
mkdir -p gpurun_out
nvidia-smi -L | head -8
nvidia-smi topo -m > gpurun_out/${P}_topo_$N.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q -s > gpurun_out/${P}_dist_$N.log 2>&1; echo "rc=$?" >> gpurun_out/${P}_dist_$N.log; tail -3 gpurun_out/${P}_dist_$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/d2h_bandwidth.py > gpurun_out/${P}_d2h_${N}gpu.json 2> gpurun_out/${P}_d2h_${N}gpu.err; cat gpurun_out/${P}_d2h_${N}gpu.json
for mode in weak strong; do
  extra=""; [ "$mode" = strong ] && extra="--no-fwd-bwd"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 30 --warmup 5 --scaling $mode $extra > gpurun_out/${P}_bench_${N}gpu_$mode.json 2> gpurun_out/${P}_bench_${N}gpu_$mode.err
  python -c "
import json
d=json.loads(open('gpurun_out/${P}_bench_${N}gpu_$mode.json').read().strip().splitlines()[-1])
print('$N GPUs $mode: value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['e2e']['mode'], 'eager', round(d['e2e']['eager']['value']), 'fwd_bwd', {k:round(v['value']) for k,v in (d.get('fwd_bwd') or {}).items()})" || tail -5 gpurun_out/${P}_bench_${N}gpu_$mode.err
done
echo done
