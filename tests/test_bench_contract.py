"""bench.py contract on a box without a GPU: the reference arm (the C port of the reference's brute-force algorithm on
the host cores -- the one place besides tests/ and smoke() that may execute oracle/) prints ONE JSON line with the
keys the driver reads; the GPU arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          cwd=ROOT, timeout=600)


def test_reference_arm_prints_the_contract_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "images/s"
    assert d["metric"] == "batched images/sec (84x84 Brax scenes)" and d["value"] > 0 and d["steps"] == 1
    assert d["config"]["workload"].startswith("configs[1]") and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_gpu_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a GPU is present: the GPU arm would run")
    p = _run("--steps", "1", "--warmup", "1", "--no-cpu", "--no-fwd-bwd", "--no-secondary")
    assert p.returncode != 0 and "needs a GPU" in (p.stderr + p.stdout)
