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


def _render_worker(rank: int, world: int, port: int, out):
    """One rank of a data-parallel differentiable render: its shard of the batch through the (CPU) oracle, the
    gradients of the SHARED parameters all-reduced over gloo -- the host logic the GPU ranks run over NCCL."""
    from types import SimpleNamespace as NS

    from oracle import jr_oracle as O
    from tests.helpers import random_mesh_scene

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = random_mesh_scene(2, n_tri=12, W=16, H=12, tex=4)
    B = 3
    pos = torch.stack([s.pos + 0.03 * b for b in range(B)])
    tex = s.texture.clone().requires_grad_(True)
    lcol = s.light.colour.clone().requires_grad_(True)
    a, b_ = shard_range(B, rank, world)
    total = torch.zeros(())
    for b in range(a, b_):
        ex = NS(position=pos[b], normal=s.nrm, uv=s.uv_texel, light=NS(direction=s.light.direction, colour=lcol), texture=tex)
        r = O.render(s.cam, "gouraud_texture", torch.ones(16, 12), (torch.zeros(16, 12, 3),), s.faces, ex)
        total = total + r.targets[0].sum()
    total.backward()
    g = [tex.grad if tex.grad is not None else torch.zeros_like(tex), lcol.grad if lcol.grad is not None else torch.zeros(3)]
    all_reduce_shared_grads(g)
    out[rank] = (g[0].clone(), g[1].clone(), b_ - a)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_sharded_render_gradients_equal_single_process():
    from types import SimpleNamespace as NS

    from oracle import jr_oracle as O
    from tests.helpers import random_mesh_scene

    world = 2
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_render_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    s = random_mesh_scene(2, n_tri=12, W=16, H=12, tex=4)
    B = 3
    tex = s.texture.clone().requires_grad_(True)
    lcol = s.light.colour.clone().requires_grad_(True)
    total = torch.zeros(())
    for b in range(B):
        ex = NS(position=s.pos + 0.03 * b, normal=s.nrm, uv=s.uv_texel, light=NS(direction=s.light.direction, colour=lcol),
                texture=tex)
        total = total + O.render(s.cam, "gouraud_texture", torch.ones(16, 12), (torch.zeros(16, 12, 3),), s.faces, ex).targets[0].sum()
    total.backward()
    assert out[0][2] + out[1][2] == B
    for r in (0, 1):    # every rank holds the full-batch gradient after the all-reduce
        torch.testing.assert_close(out[r][0], tex.grad, rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(out[r][1], lcol.grad, rtol=1e-6, atol=1e-7)
    assert float(tex.grad.abs().sum()) > 0
