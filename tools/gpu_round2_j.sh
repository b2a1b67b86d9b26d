#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for c in 1 2 4 8 16; do
  JR_E2E_CHUNKS=$c timeout 300 python bench.py --steps 20 --no-cpu --no-fwd-bwd --no-secondary > gpurun_out/r2j_c$c.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r2j_c$c.json'));print('chunks $c e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
done
python - <<'PY'
import torch, time
a=torch.empty(115605504//4, device='cuda'); h=torch.empty(115605504//4).pin_memory()
for _ in range(3): h.copy_(a, non_blocking=True); torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): h.copy_(a, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
print('pinned D2H 115.6 MB: %.3f ms = %.1f GB/s'%(ms, 115.6/ms))
e0.record()
for _ in range(10): a.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
print('pinned H2D 115.6 MB: %.3f ms = %.1f GB/s'%(ms, 115.6/ms))
PY
for v in default k4; do
  if [ "$v" = default ]; then unset JR_B200_LIB; else export JR_B200_LIB=$PWD/jaxrenderer_b200/lib/alt_$v.so; fi
  timeout 600 python bench.py --steps 5 --no-cpu --no-fwd-bwd > gpurun_out/r2j_sec_$v.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/r2j_sec_$v.json'))
print('$v', {k:round(v.get('ms_per_step',v.get('ms_per_image_eager',0)),3) for k,v in d['secondary'].items()})"
done
