"""One rank of `tests/test_gpu_distributed.py` (launched with torch.distributed.run, one process per GPU): renders its
shard of a batch with the shadow pass, back-propagates, all-reduces the gradients of the SHARED parameters over NCCL
and (rank 0) stores them."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import jaxrenderer_b200 as jr  # noqa: E402
from jaxrenderer_b200 import synthetic  # noqa: E402
from jaxrenderer_b200.distributed import all_reduce_shared_grads, shard_range  # noqa: E402


def scene(B, W, H, n_caps, dev):
    sc = synthetic.brax_like_batch(B, n_capsules=n_caps, env0=4242, with_attributes=True)
    cam = synthetic.brax_cameras(sc["eye"], sc["target"], W, H)
    g = torch.Generator().manual_seed(3)
    return sc, cam, torch.rand(B, W, H, 3, generator=g)


def grads_of_shard(sc, cam, target, a, b, n_caps, W, H, dev):
    sl = slice(a, b)
    sub = dict(sc)
    for k in ("position", "normal", "faces", "eye", "target"):
        sub[k] = sc[k][sl]
    model = synthetic.merged_model_from_batch(sub, n_caps, dev)
    camd = type(cam)(*[(t[sl] if t.ndim == 3 else t).to(dev) for t in cam])
    atlas = model.diffuse_map.clone().requires_grad_(True)
    ldir = torch.tensor((0.57735, -0.57735, 0.57735), device=dev, requires_grad=True)
    amb = torch.tensor((0.8, 0.8, 0.8), device=dev, requires_grad=True)
    light = jr.LightParameters(direction=ldir, ambient=amb, diffuse=(0.8,) * 3, specular=(0.6,) * 3)
    out = jr.Renderer.render(model._replace(diffuse_map=atlas), light, camd,
                             jr.Renderer.create_buffers(W, H, batch=b - a, device=dev),
                             shadow_param=jr.ShadowParameters(centre=sub["target"].to(dev)))
    ((out.targets[0] - target[sl].to(dev)) ** 2).sum().backward()
    return [atlas.grad, ldir.grad, amb.grad]


def main():
    out_path, B, W, H, n_caps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sc, cam, target = scene(B, W, H, n_caps, dev)
    a, b = shard_range(B, rank, world)
    g = grads_of_shard(sc, cam, target, a, b, n_caps, W, H, dev)
    all_reduce_shared_grads(g)
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({"atlas": g[0].cpu(), "ldir": g[1].cpu(), "amb": g[2].cpu(), "world": world}, out_path)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
