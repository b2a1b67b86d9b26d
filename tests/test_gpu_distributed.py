"""Multi-GPU correctness ON GPUs: N ranks each render their shard of a batch (phong_reflection_shadow
with the shadow pass), back-propagate and all-reduce the gradients of the shared scene parameters over NCCL; the result
must equal the gradients one rank computes on the whole batch.  Skipped on boxes with fewer than 2 GPUs."""
import os
import socket
import subprocess
import sys
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_nccl_sharded_gradients_equal_single_rank():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _dist_gpu_worker as Wk

    B, W, H, n_caps = 12, 84, 84, 3
    world = min(torch.cuda.device_count(), 4)
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "grads.pt")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
               os.path.join(ROOT, "tests", "_dist_gpu_worker.py"), out, str(B), str(W), str(H), str(n_caps)]
        subprocess.run(cmd, check=True, timeout=600, cwd=ROOT)
        got = torch.load(out)
    dev = torch.device("cuda", 0)
    sc, cam, target = Wk.scene(B, W, H, n_caps, dev)
    want = Wk.grads_of_shard(sc, cam, target, 0, B, n_caps, W, H, dev)
    assert got["world"] == world
    for name, g, w in (("atlas", got["atlas"], want[0]), ("light.direction", got["ldir"], want[1]), ("ambient", got["amb"], want[2])):
        w = w.cpu()
        scale = float(w.abs().max())
        err = float((g - w).abs().max())
        print(f"  {world}-rank all-reduced grad {name:16s} max|ref| {scale:.4g}  max abs err {err:.3g}  rel {err / scale:.3g}")
        assert err <= 1e-6 * scale + 1e-12, (name, err, scale)
