#!/usr/bin/env python
"""A 64x48 triangle soup with MORE THAN HALF of the canvas covered, rendered by the UNMODIFIED reference through all
seven built-in shaders (and its shadow pass) -> `tests/golden/reference_run_large.npz`.

The fixtures of `gen_reference_fixtures.py` are 28x20 at 19 % coverage: they pin conventions, not coverage (VERDICT r1,
weak 1).  Same scene generator and the same `run_soup` as there, with the triangles 2.6 x larger and the canvas 5.5 x
larger: ~1800 covered pixels per shader instead of ~100, dozens of overlapping triangles per pixel.  The stand-in's
outermost `vmap` is split over forked workers (JAX_SHIM_PROCS), results identical to the plain loop.

  JAX_SHIM_PROCS=8 python tools/gen_reference_fixtures_large.py [/root/reference]      # ~3 minutes on 8 cores

`... tiled` renders a second soup at 136x96 -- more pixels than the single-tile kernel takes (12 288), hence 3 x 2 of the 64x64 tiles of the binned CUDA path, the last column 8 pixels wide, so that the two-level
kernels (bitmasks, triangle queue, span raster, CTA-wide sweep) are pinned against the reference's own output as well
-> `tests/golden/reference_run_tiled.npz` (~6 minutes on 8 cores).

  JAX_SHIM_PROCS=8 python tools/gen_reference_fixtures_large.py /root/reference tiled
"""
from __future__ import annotations

import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_reference_fixtures as G  # noqa: E402
import numpy as np  # noqa: E402

TILED = len(sys.argv) > 2 and sys.argv[2] == "tiled"
SEED, N_TRI, W, H, SPREAD = (6, 40, 136, 96, 2.6) if TILED else (5, 48, 64, 48, 2.6)


def big_soup(seed, n_tri, W, H, tex=8):
    """`G.soup` with every triangle scaled by SPREAD about its own centre (positions only; attributes untouched)."""
    s = small(seed, n_tri, W, H, tex)
    tri = s["pos"].reshape(n_tri, 3, 3)
    centre = tri.mean(axis=1, keepdims=True)
    s["pos"] = (centre + (tri - centre) * np.float32(SPREAD)).astype(np.float32).reshape(-1, 3)
    return s


def main():
    global small
    t0 = time.time()
    small, G.soup = G.soup, big_soup
    G.run_soup(SEED, n_tri=N_TRI, W=W, H=H)
    cov = float((G.OUT[f"soup{SEED}/depth/zbuffer"] != 1.0).mean())
    assert cov > 0.5, cov
    dst = os.path.join(G.ROOT, "tests", "golden", "reference_run_tiled.npz" if TILED else "reference_run_large.npz")
    np.savez_compressed(dst, **G.OUT)
    print(f"wrote {dst}: {len(G.OUT)} arrays, {os.path.getsize(dst) / 1024:.0f} KB, coverage {cov:.2f}, {time.time() - t0:.0f}s")


if __name__ == "__main__":
    main()
