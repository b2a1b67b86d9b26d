#!/usr/bin/env python
"""Benchmark of the rasterisation hot path (contract: see the task prompt / DESIGN.md).

Workload (BASELINE.json configs[1], the config the headline metric is quoted
on): DepthShader, 84x84, batch 4096 synthetic Brax "ant-like" scenes
(ground cube + 10 capsules = 1932 triangles / 5784 vertices each, full-view
camera; the robot's body is fixed, its pose differs per environment, as in
Brax), PER GPU (weak scaling: the batch axis is sharded, no data-path
collective; `--scaling strong` fixes the GLOBAL batch at 4096 instead).  One
"step" = one `pipeline.render` of the whole per-GPU batch.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling weak|strong]

`value`   images/s with inputs resident in HBM: merged world-space positions (B, V, 3) + faces (B, T, 3), the
          arrays the reference's `pipeline.render` takes (device-timed, max over ranks).
`e2e`     images/s through the public API from HOST buffers holding what Brax produces per step -- the objects'
          transforms and the camera parameters: H2D of those, `merge_objects` (kept factored: the kernels instance
          the geometry), camera construction, `pipeline.render`, D2H of the z-buffers, all inside the timed region.
`roofline` HBM roofline of the dominant kernel (jr::k_vis3<true,true,false>), its achieved FP32 rate from an
          in-kernel N_test counter, and `roofline.issue`, the issue-slot figure.
`secondary` configs[0], [2], [3], [4]-forward of BASELINE.json: ms per step and images/s (`--no-secondary` skips).
`fwd_bwd` secondary lines: forward + backward images/s (phong_reflection_shadow, 84x84 x 4096 and 480x270 x 512).
`cpu_baseline` / `--impl reference`: the reference's brute-force algorithm
          (C port, oracle/jr_oracle_c.c) on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W = H = 84
BATCH = 4096
N_CAPSULES = 10
METRIC = "batched images/sec (84x84 Brax scenes)"
# kind "port": a vectorised, multi-threaded C restatement of the reference's algorithm, NOT the reference itself
PORT_NOTE = ("C port (AVX2 auto-vectorised, pthreads over rows) of the reference's brute-force W*H*T algorithm: faster "
             "than the reference's own JAX program would run on CPU jaxlib, so ratios against it are conservative")
UNIT = "images/s"


def _config(n_gpus: int, batch: int) -> dict:
    from jaxrenderer_b200 import synthetic

    nv, t = synthetic.scene_sizes(N_CAPSULES)
    return {
        "workload": "configs[1]: DepthShader 84x84, synthetic Brax ant-like scenes "
                    f"({t} triangles / {nv} vertices), batch {batch} per GPU",
        "shader": "depth", "width": W, "height": H, "triangles": t, "vertices": nv,
        "batch_per_gpu": batch, "global_batch": batch * n_gpus, "parallelism": f"dp{n_gpus}",
        "l2": "inputs larger than L2 (no flush needed): %.0f MB geometry per step per GPU"
              % ((nv * 12 + t * 12) * batch / 1e6),
    }


# ----------------------------------------------------------------------------- CPU arm
def _scene_host(n_images: int, env0: int = 0):
    """Merged world-space positions / faces / cameras of n_images bench scenes on the host."""
    import jaxrenderer_b200 as jr
    from jaxrenderer_b200 import synthetic

    objs, eye, tgt = synthetic.brax_like_objects(n_images, n_capsules=N_CAPSULES, env0=env0)
    m = jr.merge_objects(objs)
    cam = synthetic.brax_cameras(eye, tgt, W, H)
    return m.verts.contiguous(), m.faces, cam


def _cpu_sample(n_images: int, threads: int = 0, env0: int = 0):
    """Time the C oracle (reference algorithm, brute force) on n_images scenes."""
    import numpy as np

    from oracle import c_oracle

    pos, faces, cam = _scene_host(n_images, env0)
    z0 = np.ones((n_images, W, H), np.float32)
    args = (cam.world_to_clip.numpy(), cam.viewport.numpy(), pos.numpy(), faces.numpy(), z0)
    t0 = time.perf_counter()
    c_oracle.render_depth(*args, num_threads=threads)
    return time.perf_counter() - t0


def cpu_baseline(target_seconds: float = 12.0) -> dict:
    from oracle import c_oracle

    cores = c_oracle.max_threads()
    probe = _cpu_sample(max(2, min(cores // 8, 8)))
    per_image = probe / max(2, min(cores // 8, 8))
    n = int(max(8, min(1024, target_seconds / max(per_image, 1e-6))))
    dt = _cpu_sample(n, env0=100000)
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} images of the bench workload (brute force W*H*T per image, "
                      f"oracle/jr_oracle_c.c, {cores} pthreads), {dt:.1f} s",
            "note": PORT_NOTE}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle

    cores = c_oracle.max_threads()
    n = 48 if cores >= 32 else 16
    for _ in range(args.warmup):
        _cpu_sample(min(n, 8))
    t0 = time.perf_counter()
    for k in range(args.steps):
        _cpu_sample(n, env0=1000 * k)
    dt = (time.perf_counter() - t0) / args.steps
    value = n / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(args.gpus, BATCH),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} images per step (bounded sample of the {BATCH}-image workload), "
                                   "reference brute-force algorithm restated in C (jax is not installable here)",
                         "note": PORT_NOTE},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- GPU arm
def _csrc_sha16() -> str:
    """Hash of the kernel sources: ties `profiles/roofline_traffic.json` (an ncu capture) to the code it was taken on."""
    import hashlib

    h = hashlib.sha256()
    d = os.path.join(ROOT, "jaxrenderer_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    import jaxrenderer_b200 as jr
    from jaxrenderer_b200 import _native, pipeline, synthetic
    from jaxrenderer_b200.shaders import DepthExtraInput, DepthShader

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (ours) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _native.load()
    strong = args.scaling == "strong"
    B = args.batch // world if strong else args.batch      # images on THIS GPU
    assert B >= 1, "global batch smaller than the number of GPUs"

    # ---- synthetic inputs: what Brax produces per step, on the host (pinned), disjoint environments per rank
    objs_h, eye_h, tgt_h = synthetic.brax_like_objects(B, n_capsules=N_CAPSULES, env0=rank * B)
    tf_h = torch.stack([o.transform for o in objs_h[1:]], dim=1).contiguous().pin_memory()   # (B, n_caps, 4, 4)
    eye_h, tgt_h = eye_h.pin_memory(), tgt_h.pin_memory()
    meshes_d = [type(o.model)(*[t.to(dev) for t in o.model]) for o in objs_h]            # the robot: resident
    ground_scale = objs_h[0].local_scaling.to(dev)

    def objects_on_device(tf):
        out = [jr.ModelObject(model=meshes_d[0], local_scaling=ground_scale)]
        for i in range(N_CAPSULES):
            out.append(jr.ModelObject(model=meshes_d[i + 1], transform=tf[:, i]))
        return out

    def camera_on_device(eye, tgt):
        return jr.Renderer.create_camera_from_parameters(jr.CameraParameters(
            viewWidth=W, viewHeight=H, hfov=58.0, vfov=58.0 * H / W, position=eye, target=tgt), device=dev)

    # resident arrays of the reference boundary: merged positions (B, V, 3), faces (B, T, 3), cameras
    model_d = jr.merge_objects(objects_on_device(tf_h.to(dev)))
    pos_d = model_d.verts.materialise().contiguous()
    faces_d = model_d.faces.unsqueeze(0).expand(B, -1, -1).contiguous()
    cam_d = camera_on_device(eye_h.to(dev), tgt_h.to(dev))
    z = torch.full((B, W, H), 1.0, device=dev)
    extra_d = DepthExtraInput(position=pos_d)

    def step_resident():
        jr.render(cam_d, DepthShader, jr.Buffers(z, ()), faces_d, extra_d, inplace=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: `value`
    # Small per-GPU batches (strong scaling: 512 images = one 55-microsecond kernel per step) are launch-bound from
    # Python, eight ranks sharing the host cores: there the step is captured once into a CUDA graph and replayed
    # (`--launch graph`; `auto` does that below 2048 images per GPU).  One replay == one `pipeline.render` call.
    launch_mode = args.launch if args.launch != "auto" else ("graph" if B < 2048 else "eager")
    step_eager = step_resident
    if launch_mode == "graph":
        for _ in range(3):
            step_eager()
        torch.cuda.synchronize()
        step_graph_obj = torch.cuda.CUDAGraph()
        l0 = _native.launch_count()
        with torch.cuda.graph(step_graph_obj):
            step_eager()
        graph_launches_per_step = _native.launch_count() - l0    # kernels the library enqueued into the graph
        step_resident = step_graph_obj.replay
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _native.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for k in range(args.steps):
        step_resident()
        ev[k + 1].record()
    barrier()
    launches = _native.launch_count() - launches0
    if launch_mode == "graph":
        launches = graph_launches_per_step * args.steps          # replays do not pass through the host counter
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps))
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = B * world / (ms_per_step / 1e3)

    # in-kernel counters of one extra (untimed) launch: N_test, survivors of the filter / exact cull
    with pipeline.visibility_stats(dev) as st:
        step_eager()
    counters = st.read()

    # ---- end to end through the public API from the host buffers Brax produces: `e2e`
    z_host = torch.empty((B, W, H), dtype=torch.float32).pin_memory()
    # The batch is streamed in chunks over two CUDA streams so the H2D copy of chunk i+1, the kernels of
    # chunk i and the D2H copy of chunk i-1 overlap (PCIe is full duplex).  Every chunk is the public API:
    # merge_objects -> create_camera_from_parameters -> pipeline.render.
    # Two ways of issuing the same calls, both timed:
    #   eager  every step calls the API chunk by chunk (host cost of one chain ~0.7 ms: few chunks, host-bound);
    #   graph  the same chunked call sequence captured ONCE into a CUDA graph (what `jax.jit` gives the reference's
    #          users; the facade is allocation- and sync-free) and replayed per step: the copies read
    #          the pinned host buffers Brax refills and write the pinned result buffer, inside the timed region.
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]

    def enqueue_e2e(n_chunks):
        Bc = B // n_chunks
        cur = torch.cuda.current_stream(dev)
        for s_ in streams:
            s_.wait_stream(cur)
        for i in range(n_chunks):
            sl = slice(i * Bc, (i + 1) * Bc)
            with torch.cuda.stream(streams[i % 2]):
                tf = tf_h[sl].to(dev, non_blocking=True)
                cam_i = camera_on_device(eye_h[sl].to(dev, non_blocking=True), tgt_h[sl].to(dev, non_blocking=True))
                m = jr.merge_objects(objects_on_device(tf))
                bufs = jr.Renderer.create_buffers(W, H, batch=Bc, device=dev)
                out = jr.render(cam_i, DepthShader, jr.Buffers(bufs.zbuffer, ()), m.faces,
                                DepthExtraInput(position=m.verts), inplace=True)
                z_host[sl].copy_(out.zbuffer, non_blocking=True)
        for s_ in streams:
            cur.wait_stream(s_)

    def pick_chunks(want):
        n = max(1, min(want, B))
        while B % n:
            n -= 1
        return n

    def time_e2e(step):
        for _ in range(3):
            step()
        barrier()
        steps = max(3, min(args.steps, 20))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()) / steps, steps

    eager_chunks = pick_chunks(int(os.environ.get("JR_E2E_CHUNKS", "0")) or (2 if B >= 64 else 1))

    def step_eager():
        enqueue_e2e(eager_chunks)
        torch.cuda.current_stream(dev).synchronize()

    eager_ms, e2e_steps = time_e2e(step_eager)
    e2e_ms, e2e_mode, e2e_chunks = eager_ms, "eager", eager_chunks
    graph_note = None
    if args.e2e != "eager":
        graph_chunks = pick_chunks(int(os.environ.get("JR_E2E_GRAPH_CHUNKS", "0")) or (8 if B >= 1024 else 2))
        try:
            enqueue_e2e(graph_chunks)                       # warm-up: workspaces, memoised constants
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                enqueue_e2e(graph_chunks)

            def step_graph():
                graph.replay()
                torch.cuda.current_stream(dev).synchronize()

            z_host.zero_()
            step_graph()
            assert torch.equal(z_host, z.cpu()), "graph-replayed e2e output differs from the device-resident output"
            graph_ms, e2e_steps = time_e2e(step_graph)
            e2e_ms, e2e_mode, e2e_chunks = graph_ms, "graph", graph_chunks
        except Exception as exc:  # capture not possible: the eager figure stands, and the line says why
            graph_note = f"{type(exc).__name__}: {exc}"[:200]
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None   # sampled over both timed regions (resident + e2e)
    h2d = (tf_h.numel() + eye_h.numel() + tgt_h.numel()) * 4
    d2h = z_host.numel() * 4
    # parity guard: the e2e result (instanced geometry) equals the resident result (merged arrays) bit for bit
    assert torch.equal(z_host, z.cpu()), "e2e output differs from the device-resident output"

    fwd_bwd = None
    if not args.no_fwd_bwd:
        # the metric's 84x84 scenes at the metric's batch, and configs[4] as named (480x270, T = 3276, B = 512)
        fwd_bwd = {"84x84": fwd_bwd_secondary(dev, rank, world, shape=(84, 84, 10, 4096)),
                   "480x270": fwd_bwd_secondary(dev, rank, world, shape=(480, 270, 17, 512))}
    secondary = None
    if not args.no_secondary and world == 1:
        secondary = secondary_configs(dev)
    if rank == 0:
        nv, ntri = synthetic.scene_sizes(N_CAPSULES)
        alg_bytes = (12 * nv + 12 * ntri + 4 * W * H) * B       # geometry read once + z write
        med_ms = per_launch_ms[len(per_launch_ms) // 2]
        avg_ms = total_ms / args.steps
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = warp_inst = None
        traffic_note = "no ncu capture on record"
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
            if prof.get("csrc_sha16") == _csrc_sha16():
                traffic = prof.get("dram_bytes_per_launch")
                warp_inst = prof.get("warp_instructions_per_launch")
                traffic_note = prof.get("source", "")
            else:
                traffic_note = ("profiles/roofline_traffic.json was captured on other kernel sources (csrc hash "
                                "differs): not reported")
        except (OSError, ValueError):
            pass
        achieved = alg_bytes / (avg_ms / 1e3) / 1e9
        sm_mhz = float(clocks["sm_mhz"]) if clocks and clocks.get("sm_mhz") else None
        # FP32: SURVEY 8d's algorithmic flops 28 V + 90 T + 16 N_test per image, N_test from the kernel's own counter
        n_test = counters["n_test"]
        flops = (28.0 * nv + 90.0 * ntri) * B + 16.0 * n_test
        fp32_peak = 148 * 128 * 2 * (sm_mhz or 1965.0) * 1e6
        fp32 = {"bound": "fp32", "achieved": flops / (avg_ms / 1e3) / 1e12, "peak": fp32_peak / 1e12, "unit": "TFLOP/s",
                "frac": flops / (avg_ms / 1e3) / fp32_peak, "n_test_per_launch": n_test,
                "n_test_reference_per_launch": W * H * ntri * B, "counters": counters,
                "peak_source": "148 SMs x 128 fp32 lanes x 2 x SM clock sampled in this run"}
        issue = None
        if warp_inst and sm_mhz:
            peak_issue = 148 * 4 * sm_mhz * 1e6
            ach_issue = warp_inst * (B / BATCH) / (avg_ms / 1e3)
            issue = {"bound": "issue slots", "achieved": ach_issue / 1e9, "peak": peak_issue / 1e9,
                     "unit": "G warp-instructions/s", "frac": ach_issue / peak_issue,
                     "warp_instructions_per_launch": warp_inst * (B / BATCH)}
        cfg = _config(world, B)
        cfg["launch"] = ("one CUDA-graph replay per step (the render call captured once)" if launch_mode == "graph"
                         else "eager (one pipeline.render call per step)")
        if strong:
            cfg["workload"] += f" (strong scaling: global batch {args.batch} split over {world} GPUs)"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "clocks": clocks,
            "e2e": {"value": B * world / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "mode": e2e_mode, "chunks": e2e_chunks,
                    "eager": {"value": B * world / (eager_ms / 1e3), "ms_per_step": eager_ms, "chunks": eager_chunks},
                    "graph_error": graph_note,
                    "inputs": "host: object transforms + camera parameters (what Brax produces); geometry instanced "
                              "in-kernel from the resident robot meshes"},
            "gpu_launches": int(launches),
            "roofline": {
                "kernel": "jr::k_vis3<true,true,false> (filter + exact triangle setup + raster + depth resolve)",
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic * (B / BATCH) if traffic else None,
                "traffic_source": traffic_note,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "algorithmic_bytes_per_launch": alg_bytes,
                "launch_ms_avg": avg_ms, "launch_ms_median": med_ms,
                "fp32": fp32, "issue": issue,
            },
        }
        if fwd_bwd is not None:
            line["fwd_bwd"] = fwd_bwd
            line["fwd_bwd_images_per_s"] = fwd_bwd["84x84"]["value"]
            line["fwd_bwd_480x270_images_per_s"] = fwd_bwd["480x270"]["value"]
        if secondary is not None:
            line["secondary"] = secondary
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _time_steps(fn, steps: int, warmup: int = 3) -> float:
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


_VIS_KERNELS = ("memset+k_setup_bin", "k_raster_tile", "k_vis3", "k_vis2")
_SHADE_KERNELS = ("memset+k_mark_visible", "k_tri_attr", "k_shade_rec", "k_shade_rec_u8", "k_shade")


def _hbm_peak() -> tuple:
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except (OSError, ValueError, KeyError):
        return 6650.0, "fallback 6650 GB/s"


def _stage_rooflines(step, B: int, W: int, H: int, n_verts: int, n_tris: int, shade_bytes_per_pixel: int,
                     reps: int = 3) -> dict:
    """Per-stage rooflines of one secondary step from the library's own per-kernel event timing
    (`jr_debug_kernel_timing`, include/jr_b200.h): the step runs `reps` more times with an event after every launch;
    kernels are grouped per entry-point call into the visibility stage (binning + raster, or the single-tile kernel)
    and the shading stage (visible-triangle lists + attribute records + pixel stage).  Algorithmic bytes
    (SURVEY 8d / DESIGN 3): visibility = geometry once + one 4-byte plane out = B (12 Nv + 12 T + 4 W H); shading =
    `shade_bytes_per_pixel` x B W H (G-buffer in, z + colour out, + the shadow-map sample) -- varyings and texels of
    the visible triangles are left out, so the shading fraction is a lower bound."""
    import torch

    from jaxrenderer_b200 import _native

    peak, peak_src = _hbm_peak()
    torch.cuda.synchronize()
    _native.kernel_timing(True)
    try:
        for _ in range(reps):
            step()
        torch.cuda.synchronize()
        rows = _native.kernel_times()
    finally:
        _native.kernel_timing(False)
    if not rows:
        return {}
    calls_per_step = (max(c for _, _, c in rows) + 1) // reps
    stages, order = {}, []
    for name, ms, call in rows:
        c = call % calls_per_step
        kind = "visibility" if name in _VIS_KERNELS else ("shading" if name in _SHADE_KERNELS else "other")
        key = (c, kind)
        if key not in stages:
            stages[key] = {}
            order.append(key)
        stages[key][name] = stages[key].get(name, 0.0) + ms / reps
    out, depth_pass_seen = [], False
    n_vis = sum(1 for _, k in order if k == "visibility")
    for c, kind in order:
        ks = stages[(c, kind)]
        ms = sum(ks.values())
        row = {"stage": kind, "call": c, "kernels_ms": {k: round(v, 5) for k, v in ks.items()}, "ms": ms}
        if kind == "visibility":
            if n_vis > 1:
                row["stage"] = "visibility (shadow pass)" if not depth_pass_seen else "visibility (main pass)"
            depth_pass_seen = True
            nbytes = B * (12 * n_verts + 12 * n_tris + 4 * W * H)
        elif kind == "shading":
            nbytes = B * W * H * shade_bytes_per_pixel
        else:
            nbytes = None
        if nbytes:
            row["roofline"] = {"bound": "hbm", "algorithmic_bytes": nbytes, "achieved": nbytes / (ms / 1e3) / 1e9,
                               "peak": peak, "unit": "GB/s", "frac": nbytes / (ms / 1e3) / 1e9 / peak}
        out.append(row)
    return {"stages": out, "kernels_ms_total": sum(r["ms"] for r in out), "peak_source": peak_src,
            "timing": f"CUDA events recorded by the library after each launch, mean of {reps} steps"}


def secondary_configs(dev, steps: int = 5) -> dict:
    """The other configurations BASELINE.json names, device-resident, one GPU (secondary lines):
    configs[0] simple_cube 640x480 latency (B = 1, both shadow modes), configs[2] gouraud_texture 32x32 x 16384,
    configs[3] phong_reflection_shadow 960x540 x 256 (19 980 triangles, shadow pass), configs[4] forward 480x270 x 512."""
    import torch

    import jaxrenderer_b200 as jr
    from jaxrenderer_b200 import synthetic
    from jaxrenderer_b200.shaders import GouraudTextureExtraInput, GouraudTextureShader

    out = {}
    light = jr.LightParameters(direction=(0.57735, -0.57735, 0.57735), ambient=(0.8,) * 3, diffuse=(0.8,) * 3,
                               specular=(0.6,) * 3)
    # ---- configs[0]: examples/simple_cube.py, one image, eager and CUDA-graph replay
    cube = jr.create_cube(torch.ones(3), torch.ones(2), torch.zeros(2, 2, 3).index_fill_(2, torch.tensor([2]), 1.0),
                          torch.ones(2, 2) * 2.0)
    objs = [jr.ModelObject(model=type(cube)(*[t.to(dev) for t in cube]))]
    camp = jr.CameraParameters(viewWidth=640, viewHeight=480, position=torch.tensor([2.0, 4.0, 1.0], device=dev))
    for name, sp in (("no_shadow", None), ("shadow", jr.ShadowParameters())):
        fn = lambda: jr.Renderer.get_camera_image(objs, jr.LightParameters(), camp, 640, 480, shadow_param=sp)  # noqa: E731
        ms = _time_steps(fn, 20)
        img, ms_g = fn(), None
        try:    # the whole facade call replayed as ONE CUDA graph (allocation- and sync-free after warm-up)
            g = torch.cuda.CUDAGraph()
            s_ = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(s_):
                fn(); fn()
                with torch.cuda.graph(g, stream=s_):
                    img = fn()
            torch.cuda.synchronize()
            ms_g = _time_steps(g.replay, 20)
        except RuntimeError as e:   # pragma: no cover
            ms_g = f"capture failed: {str(e)[:80]}"
            torch.cuda.synchronize()
        out[f"configs[0] simple_cube 640x480 B=1 {name}"] = {"ms_per_image_eager": ms, "ms_per_image_cuda_graph": ms_g,
                                                             "covered_pixels": int((img != 1.0).any(-1).sum())}
    # ---- configs[2]: gouraud_texture 32x32, 16384 images
    B, Wd, Hd = 16384, 32, 32
    sc = synthetic.brax_like_batch(B, n_capsules=10, env0=20_000_000, with_attributes=True)
    cam = synthetic.brax_cameras(sc["eye"], sc["target"], Wd, Hd)
    cam = type(cam)(*[t.to(dev) for t in cam])
    ex = GouraudTextureExtraInput(sc["position"].to(dev), sc["normal"].to(dev), (sc["uv"] * 100).to(dev),
                                  jr.LightSource(torch.tensor((0.57735, -0.57735, 0.57735), device=dev), torch.ones(3, device=dev)),
                                  synthetic.checker_texture().to(dev))
    faces = sc["faces"].to(dev)
    z, c = torch.empty(B, Wd, Hd, device=dev), torch.empty(B, Wd, Hd, 3, device=dev)

    def step3():
        z.fill_(1.0); c.fill_(0.0)
        jr.render(cam, GouraudTextureShader, jr.Buffers(z, (c,)), faces, ex, inplace=True)
    ms = _time_steps(step3, steps)
    out["configs[2] gouraud_texture 32x32 B=16384 T=1932"] = {
        "ms_per_step": ms, "images_per_s": B / ms * 1e3,
        **_stage_rooflines(step3, B, Wd, Hd, sc["position"].shape[-2], faces.shape[-2], 24)}
    del sc, ex, faces, z, c
    # ---- configs[3] (B = 256) and configs[4] forward (B = 512): Renderer.render with the shadow pass
    for key, (Wd, Hd, n_caps, B) in (("configs[3] phong_reflection_shadow 960x540 B=256 T=19980", (960, 540, 104, 256)),
                                     ("configs[4] forward phong_reflection_shadow 480x270 B=512 T=3276", (480, 270, 17, 512))):
        sc = synthetic.brax_like_batch(B, n_capsules=n_caps, env0=30_000_000, with_attributes=True)
        cam = synthetic.brax_cameras(sc["eye"], sc["target"], Wd, Hd)
        cam = type(cam)(*[t.to(dev) for t in cam])
        model = synthetic.merged_model_from_batch(sc, n_caps, dev)
        sp = jr.ShadowParameters(centre=sc["target"].to(dev))
        bufs = jr.Renderer.create_buffers(Wd, Hd, batch=B, device=dev)

        def step():
            bufs.zbuffer.fill_(1.0)
            jr.Renderer.render(model, light, cam, bufs, shadow_param=sp, inplace=True)
        ms = _time_steps(step, steps)
        out[key] = {"ms_per_step": ms, "images_per_s": B / ms * 1e3,
                    "pixels_per_s": B * Wd * Hd / ms * 1e3,
                    **_stage_rooflines(step, B, Wd, Hd, sc["position"].shape[-2], sc["faces"].shape[-2], 28)}
        del sc, model, bufs
        torch.cuda.empty_cache()
    return out


def fwd_bwd_secondary(dev, rank: int, world: int, steps: int = 5, shape=(480, 270, 17, 64)) -> dict:
    """Secondary line (BASELINE metric: "fwd+bwd images/sec"; configs[4]):
    Renderer.render with the shadow pass at 480x270, 3276 triangles, 512 images per GPU, loss =
    mean((canvas - target)^2), gradients w.r.t. light, world_to_clip and the SHARED diffuse atlas;
    the shared gradients are all-reduced across ranks (NCCL) inside the timed region."""
    import torch
    import torch.distributed as dist

    import jaxrenderer_b200 as jr
    from jaxrenderer_b200 import synthetic
    from jaxrenderer_b200.distributed import all_reduce_shared_grads

    Wd, Hd, n_caps, Bd = shape
    sc = synthetic.brax_like_batch(Bd, n_capsules=n_caps, env0=10_000_000 + rank * Bd, with_attributes=True)
    cam = synthetic.brax_cameras(sc["eye"], sc["target"], Wd, Hd)
    cam = type(cam)(*[t.to(dev) for t in cam])
    model = synthetic.merged_model_from_batch(sc, n_caps, dev)
    light0 = jr.LightParameters(direction=(0.57735, -0.57735, 0.57735), ambient=(0.8,) * 3, diffuse=(0.8,) * 3,
                                specular=(0.6,) * 3)
    sp = jr.ShadowParameters(centre=sc["target"].to(dev))
    target = torch.rand(Bd, Wd, Hd, 3, device=dev)
    atlas = model.diffuse_map.clone().requires_grad_(True)
    ldir = torch.tensor(light0.direction, device=dev, requires_grad=True)
    amb = torch.tensor(light0.ambient, device=dev, requires_grad=True)
    w2c = cam.world_to_clip.clone().requires_grad_(True)

    def step():
        for p in (atlas, ldir, amb, w2c):
            p.grad = None
        out = jr.Renderer.render(model._replace(diffuse_map=atlas), light0._replace(direction=ldir, ambient=amb),
                                 cam._replace(world_to_clip=w2c), jr.Renderer.create_buffers(Wd, Hd, batch=Bd, device=dev),
                                 shadow_param=sp)
        ((out.targets[0] - target) ** 2).mean().backward()
        all_reduce_shared_grads([atlas.grad, ldir.grad, amb.grad])   # w2c is per image: no exchange

    for _ in range(3):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    kernels_ms = None
    if world == 1:   # where the step goes, kernel by kernel (the library's own event timing, one more step)
        from jaxrenderer_b200 import _native
        _native.kernel_timing(True)
        try:
            step()
            torch.cuda.synchronize()
            kernels_ms = {}
            for name, k_ms, _ in _native.kernel_times():
                kernels_ms[name] = round(kernels_ms.get(name, 0.0) + k_ms, 5)
        finally:
            _native.kernel_timing(False)
    return {"value": Bd * world / (ms / 1e3), "unit": "images/s (forward + backward)", "ms_per_step": ms,
            "steps": steps, "kernels_ms": kernels_ms,
            "config": {"workload": f"phong_reflection_shadow {Wd}x{Hd}, {synthetic.scene_sizes(n_caps)[1]} triangles, "
                                   f"{Bd} images per GPU, grads w.r.t. light, world_to_clip, shared diffuse atlas",
                       "allreduce_floats": int(atlas.numel() + 6) if world > 1 else 0}}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--batch", type=int, default=BATCH, help="images per GPU (weak) / in total (strong)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-fwd-bwd", action="store_true", help="skip the secondary forward+backward measurement")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary configurations (configs[0], [2], [3], [4])")
    ap.add_argument("--launch", choices=("auto", "graph", "eager"), default="auto",
                    help="how the device-resident step is issued: eagerly, as one CUDA-graph replay, or auto "
                         "(graph below 2048 images per GPU, where the Python launch path is the bottleneck)")
    ap.add_argument("--e2e", choices=("graph", "eager"), default="graph",
                    help="how the e2e step issues the public-API calls: captured once into a CUDA graph and replayed "
                         "(default; the eager figure is reported beside it) or eagerly every step")
    ap.add_argument("--scaling", choices=("weak", "strong"), default="weak",
                    help="weak: --batch images per GPU; strong: --batch images in total, split over the GPUs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
