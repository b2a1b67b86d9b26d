#!/usr/bin/env bash
# multi-GPU pass: NCCL gradient equivalence + weak / strong scaling of the bench
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q -s > gpurun_out/r2k_dist_$N.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_dist_$N.log; tail -6 gpurun_out/r2k_dist_$N.log
for mode in weak strong; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 30 --warmup 5 --scaling $mode > gpurun_out/r2k_bench_${N}gpu_$mode.json 2> gpurun_out/r2k_bench_${N}gpu_$mode.err
  python -c "
import json
d=json.load(open('gpurun_out/r2k_bench_${N}gpu_$mode.json'))
print('$N GPUs $mode: value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'fwd_bwd', {k:round(v['value']) for k,v in d.get('fwd_bwd',{}).items()})" || tail -5 gpurun_out/r2k_bench_${N}gpu_$mode.err
done
