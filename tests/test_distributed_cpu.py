"""World-size-2 gloo tests of the N>1 host logic (batch sharding, shared-gradient all-reduce)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jaxrenderer_b200.distributed import all_reduce_shared_grads, shard_batch, shard_range


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch = torch.arange(7 * 3, dtype=torch.float32).reshape(7, 3)
    mine = shard_batch(batch, rank, world)
    # per-rank "gradient" of shared parameters = sum over the local shard
    g_light = mine.sum(0)
    g_tex = (mine ** 2).sum().reshape(1, 1).expand(2, 2).contiguous()
    all_reduce_shared_grads([g_light, g_tex])
    ok = torch.allclose(g_light, batch.sum(0)) and torch.allclose(g_tex, (batch ** 2).sum().expand(2, 2))
    out[rank] = (bool(ok), mine.shape[0])
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_the_batch():
    for n in (1, 7, 8, 4096):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_gloo_world2_shared_grad_allreduce():
    world = 2
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0] == (True, 4) and out[1] == (True, 3)
