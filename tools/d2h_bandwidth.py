#!/usr/bin/env python
"""Host-link ceiling of the `e2e` leg: pinned device->host (and host->device) copy bandwidth of every rank of a node
running CONCURRENTLY, with the bench's own result size (4096 x 84 x 84 float32 = 115.6 MB per rank and step).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29519 tools/d2h_bandwidth.py

Rank 0 prints one JSON line: per-rank GB/s (min / max) and the node aggregate.  The e2e figure of bench.py cannot exceed
aggregate / 28 224 bytes per image."""
import json
import os

import torch
import torch.distributed as dist


def main() -> None:
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = 4096 * 84 * 84
    d = torch.empty(n, device=dev)
    h = torch.empty(n).pin_memory()
    out = {}
    for name, copy in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
        for _ in range(3):
            copy()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 20
        e0.record()
        for _ in range(steps):
            copy()
        e1.record()
        torch.cuda.synchronize()
        gbs = n * 4 * steps / (e0.elapsed_time(e1) / 1e3) / 1e9
        t = torch.tensor([gbs], device=dev, dtype=torch.float64)
        if world > 1:
            allv = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allv, t)
            vals = [float(v.item()) for v in allv]
        else:
            vals = [gbs]
        out[name] = {"per_rank_gbs_min": min(vals), "per_rank_gbs_max": max(vals), "aggregate_gbs": sum(vals)}
    if int(os.environ.get("RANK", "0")) == 0:
        out["n_gpus"] = world
        out["bytes_per_copy"] = n * 4
        out["e2e_ceiling_images_per_s"] = out["d2h"]["aggregate_gbs"] * 1e9 / (84 * 84 * 4)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
