// jr_visibility.cuh -- k_vis2: the tuned visibility kernel.
//
// One CTA (256 threads) per (image, screen tile); the tile's packed 64-bit
// (orderable z | triangle id) keys live in shared memory.  There is NO block
// barrier inside the triangle loop (ncu showed barrier stalls dominating a
// queue-per-round variant); work is balanced at warp granularity instead:
//
//   lane = triangle   gather vertices, clip x/y/w (z only for survivors), det,
//                     back-face / degenerate / behind-camera cull, approximate
//                     bbox (+-0.5 px), exact LU inverse for survivors.
//   small  (<= 16 px) rasterised by the owning lane (threshold swept on B200: 8/16/32/64 px).
//   medium (<= 1024)  rasterised by the whole warp right away: the owner's
//                     record is broadcast with shuffles, lanes cover the bbox
//                     pixels flat (no integer division).
//   large  (> 1024)   queued in shared memory, rasterised by the whole CTA after
//                     the loop (flat; the ground plane of Brax scenes).
//   resolve           keys -> tri_id (+ z for the depth shader).
//
// The 64-bit shared atomicMin is an explicit ld.shared / atom.shared.cas.b64 loop on
// a 32-bit shared-window address (the compiler's generic-pointer emulation costs ~3x).
//
// K32 variant (depth shader, no triangle-id output: the bench workload and every shadow pass): the
// key is the orderable z alone -> half the shared memory, so 5 CTAs (40 warps) fit an SM instead
// of 3.  The kernel is issue/latency bound and the extra warps are what pays (measured on B200,
// 84x84 x 4096: 0.484 ms with 64-bit keys, 0.506 with 32-bit keys at the same 3 CTAs/SM, 0.442 at
// 4, 0.436 at 5 with 76 B of register spill, 0.454 at 6).
#pragma once
#include <type_traits>
#include "jr_device.cuh"

namespace jr {

constexpr int V2_THREADS = 256;
#ifndef JR_K32_THREADS
#define JR_K32_THREADS 256
#endif
constexpr int V2_K32_THREADS = JR_K32_THREADS;  // CTA size of the z-only-key variant
constexpr int V2_BIGCAP = 32;
#ifndef JR_SMALL_AREA
#define JR_SMALL_AREA 16
#endif
#ifndef JR_MEDIUM_AREA
#define JR_MEDIUM_AREA 1024
#endif
#ifndef JR_K32_CTAS
#define JR_K32_CTAS 5
#endif
#ifndef JR_K64_CTAS
#define JR_K64_CTAS 4
#endif
constexpr int V2_K32_CTAS = JR_K32_CTAS;        // resident CTAs per SM targeted by the z-only-key variant
constexpr int V2_SMALL_AREA = JR_SMALL_AREA;    // boxes up to this many pixels: the owning lane
constexpr int V2_MEDIUM_AREA = JR_MEDIUM_AREA;  // up to this: the owning warp; above: the whole CTA
constexpr int V2_HIER_AREA = 256;   // warp-cooperative boxes from this size use the hierarchical raster

struct V2Layout { size_t keys, xs, ys, bigq, total; };
// key_bytes: 8 for packed (z | triangle id) keys, 4 for the z-only keys of the depth shader
__host__ __device__ inline V2Layout v2_layout(int tile_w, int tile_h, int key_bytes) {
  V2Layout L;
  L.keys = 0;
  L.xs = ((size_t)tile_w * tile_h * key_bytes + 15) & ~(size_t)15;
  L.ys = L.xs + (size_t)tile_w * 4;
  size_t e = L.ys + (size_t)tile_h * 4;
  L.bigq = (e + 15) & ~(size_t)15;
  L.total = L.bigq + (size_t)V2_BIGCAP * 64;
  return L;
}

__device__ __forceinline__ void key_min(uint32_t saddr, unsigned long long key) {
  unsigned long long old;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(old) : "r"(saddr));
  while (key < old) {
    unsigned long long prev;
    asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(prev) : "r"(saddr), "l"(old), "l"(key) : "memory");
    if (prev == old) break;
    old = prev;
  }
}

// One fragment into the tile's key buffer.  K32 (depth shader without a triangle-id output): the key is
// the orderable depth alone and the update is ONE native shared-memory atomic min; equal depths need no
// tie-break because only z is written.  0xFFFFFFFF is the "empty" mark, so the canonical NaN (whose
// orderable image it is) is stored as 0xFFFFFFFE -- still a NaN after from_orderable.
template <bool K32>
__device__ __forceinline__ void put_key(uint32_t keys_saddr, int idx, float zw, unsigned tri) {
  if (K32) {
    const uint32_t k = min(orderable(zw), 0xFFFFFFFEu);
    const uint32_t addr = keys_saddr + (uint32_t)idx * 4u;
    uint32_t old;  // plain load first: occluded fragments (most of them) never reach the atomic unit
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(old) : "r"(addr));
    if (k < old) asm volatile("red.shared.min.u32 [%0], %1;" ::"r"(addr), "r"(k) : "memory");
  } else {
    key_min(keys_saddr + (uint32_t)idx * 8u, ((unsigned long long)orderable(zw) << 32) | tri);
  }
}

// Warp-cooperative, EXACT hierarchical rasterisation of one triangle's bbox (tile-local, inclusive).
// Every rounded fp32 op is monotone, so fl(fl(xn*i0 + yn*i1) + i2) is monotone in xn and in yn and,
// over a block of pixels, attains its max / min at one of the 4 block corners: a 4x8 block whose
// corner max is < 0 for some edge has no inside pixel (skipped); one whose corner min is >= 0 for
// all edges is fully inside (edge tests skipped).  One lane classifies one block (32 blocks per
// round); surviving blocks get one lane per pixel.  Pays off from ~8 blocks (256 px) upwards.
template <bool K32>
__device__ __forceinline__ void raster_hier_warp(const float* inv, const float* zc, unsigned tri, int x0, int y0,
                                                 int x1, int y1, int lane, const float* xs, const float* ys,
                                                 uint32_t keys_saddr, int key_stride, float vp22, float vp23) {
  const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
  const int nby = (bh + 7) >> 3;
  const int nblk = ((bw + 3) >> 2) * nby;
  const float rnby = 1.0f / (float)nby;
  for (int base = 0; base < nblk; base += 32) {
    const int blk = base + lane;
    bool live = false, full = false;
    int xa = 0, ya = 0;
    if (blk < nblk) {
      const int bx = (int)(((float)blk + 0.5f) * rnby), by = blk - bx * nby;
      xa = x0 + 4 * bx; ya = y0 + 8 * by;
      const int xb = min(xa + 3, x1), yb = min(ya + 7, y1);
      const float xna = xs[xa], xnb = xs[xb], yna = ys[ya], ynb = ys[yb];
      live = true; full = true;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float pa = xna * inv[k], pb = xnb * inv[k];
        const float qa = yna * inv[3 + k], qb = ynb * inv[3 + k];
        const float v0 = (pa + qa) + inv[6 + k], v1 = (pa + qb) + inv[6 + k];
        const float v2 = (pb + qa) + inv[6 + k], v3 = (pb + qb) + inv[6 + k];
        live = live && (fmaxf(fmaxf(v0, v1), fmaxf(v2, v3)) >= 0.f);
        full = full && (v0 >= 0.f) && (v1 >= 0.f) && (v2 >= 0.f) && (v3 >= 0.f);
      }
    }
    unsigned m_live = __ballot_sync(0xffffffffu, live);
    const unsigned m_full = __ballot_sync(0xffffffffu, full);
    const int xy = (xa << 16) | ya;
    while (m_live) {
      const int j = __ffs(m_live) - 1;
      m_live &= m_live - 1;
      const int xyj = __shfl_sync(0xffffffffu, xy, j);
      const int x = (xyj >> 16) + (lane >> 3), y = (xyj & 0xffff) + (lane & 7);
      if (x <= x1 && y <= y1) {
        const float xn = xs[x], yn = ys[y];
        const float c0 = (xn * inv[0] + yn * inv[3]) + inv[6];
        const float c1 = (xn * inv[1] + yn * inv[4]) + inv[7];
        const float c2 = (xn * inv[2] + yn * inv[5]) + inv[8];
        if (((m_full >> j) & 1u) || (c0 >= 0.f && c1 >= 0.f && c2 >= 0.f)) {
          const float z = (c0 * zc[0] + c1 * zc[1]) + c2 * zc[2];
          const float zw = z * vp22 + vp23;
          put_key<K32>(keys_saddr, x * key_stride + y, zw, tri);
        }
      }
    }
  }
}

// Warp-cooperative, EXACT span rasterisation of one triangle's bbox (tile-local, inclusive): for boxes that are
// mostly empty -- the long thin triangles of a capsule's cylinder on a large canvas fill ~7 % of their box, and
// there are a thousand of them per 960x540 humanoid frame.  A lane takes one LINE of the box along its longer side
// (a column when the box is wider than tall, else a row) and finds the interval of the line that is inside: for a
// fixed line every edge function  fl(fl(p + fl(v * i)) + c)  is monotone in the walked coordinate (each rounded op
// is, and the pixel -> NDC table is), so "edge k >= 0" holds on a prefix or a suffix of the line.  The switch point
// is located from the analytic root (approximate arithmetic) and walked to the exact place with the rasterisers' own
// expression; the pixels of the interval are then evaluated -- and TESTED -- exactly like everywhere else, so the
// interval only has to contain the inside set for the result to be bit-identical (it is exact, the test is a guard).
// fp32 addition and multiplication commute, so one expression serves both orientations.
template <bool K32>
__device__ __forceinline__ void raster_span_warp(const float* inv, const float* zc, unsigned tri, int x0, int y0,
                                                 int x1, int y1, int lane, const float* xs, const float* ys,
                                                 uint32_t keys_saddr, int key_stride, float vp22, float vp23) {
  const bool cols = (x1 - x0) >= (y1 - y0);          // lanes = columns, walk along y
  const int u0 = cols ? x0 : y0, u1 = cols ? x1 : y1, v0 = cols ? y0 : x0, v1 = cols ? y1 : x1;
  const float* __restrict__ us = cols ? xs : ys;
  const float* __restrict__ vs = cols ? ys : xs;
  float iu[3], iv[3];                                // rows of inv multiplying the lane's / the walked coordinate
#pragma unroll
  for (int k = 0; k < 3; ++k) { iu[k] = cols ? inv[k] : inv[3 + k]; iv[k] = cols ? inv[3 + k] : inv[k]; }
  const float vs0 = vs[v0];
  const float idx_per_v = (v1 > v0) ? (float)(v1 - v0) * __fdividef(1.f, vs[v1] - vs0) : 0.f;  // table index per NDC unit
  for (int u = u0 + lane; u <= u1; u += 32) {
    const float un = us[u];
    int lo = v0, hi = v1;
    bool empty = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (empty) continue;
      const float pk = un * iu[k], ik = iv[k], ck = inv[6 + k];
      const bool in_lo = ((pk + vs[lo] * ik) + ck) >= 0.f;
      const bool in_hi = ((pk + vs[hi] * ik) + ck) >= 0.f;
      if (in_lo && in_hi) continue;
      if (!in_lo && !in_hi) { empty = true; continue; }
      const float root = (float)v0 + (-(pk + ck) * __fdividef(1.f, ik) - vs0) * idx_per_v;
      int l = min(max((int)floorf(fminf(fmaxf(root, (float)lo), (float)hi)), lo), hi - 1);
      while (l > lo && (((pk + vs[l] * ik) + ck) >= 0.f) != in_lo) --l;
      while (l + 1 < hi && (((pk + vs[l + 1] * ik) + ck) >= 0.f) == in_lo) ++l;
      if (in_lo) hi = l; else lo = l + 1;
    }
    if (empty) continue;
    for (int v = lo; v <= hi; ++v) {
      const int x = cols ? u : v, y = cols ? v : u;
      const float xn = xs[x], yn = ys[y];
      const float c0 = (xn * inv[0] + yn * inv[3]) + inv[6];
      const float c1 = (xn * inv[1] + yn * inv[4]) + inv[7];
      const float c2 = (xn * inv[2] + yn * inv[5]) + inv[8];
      if (c0 >= 0.f && c1 >= 0.f && c2 >= 0.f) {
        const float z = (c0 * zc[0] + c1 * zc[1]) + c2 * zc[2];
        const float zw = z * vp22 + vp23;
        put_key<K32>(keys_saddr, x * key_stride + y, zw, tri);
      }
    }
  }
}

// The same, with the interval of a line taken from the ANALYTIC roots of the three edge functions, widened by a bound
// on everything that separates them from the rounded evaluation, instead of an exact search (which costs ~50
// instructions per edge and line).  The per-pixel test below is the rasterisers' own, so the result is bit-identical
// as long as the interval CONTAINS every accepted pixel.  Bound, in table-index units, for edge k on the line
// (p = fl(u * i_u), i = i_v, c):  the rounded value  fl(fl(p + fl(v i)) + c)  differs from the real  p + v i + c  by at
// most E = 3 * 2^-24 (|p| + |i| max|v| + |c|), so an accepted pixel satisfies  v i >= -(p + c) - E, i.e. lies within
// E / |i| (x index-per-NDC) of the real root on its inner side; the root itself is evaluated in fp32 with a relative
// error below 2^-20 (reciprocal approximation, three roundings) and the index <-> NDC table is linear to 10^-3 pixel.
// Margin used: E-term + |root| 2^-18 + 0.05.  A vanishing i gives a NaN / infinite root, which fmaxf / fminf ignore
// (no constraint from that edge: every pixel of the line is tested).
template <bool K32>
__device__ __forceinline__ void raster_span2_warp(const float* inv, const float* zc, unsigned tri, int x0, int y0,
                                                  int x1, int y1, int lane, const float* xs, const float* ys,
                                                  uint32_t keys_saddr, int key_stride, float vp22, float vp23) {
  const bool cols = (x1 - x0) >= (y1 - y0);          // lanes = columns, walk along y
  const int u0 = cols ? x0 : y0, u1 = cols ? x1 : y1, v0 = cols ? y0 : x0, v1 = cols ? y1 : x1;
  const float* __restrict__ us = cols ? xs : ys;
  const float* __restrict__ vs = cols ? ys : xs;
  const float vs0 = vs[v0], vs1 = vs[v1];
  const float idx_per_v = (v1 > v0) ? (float)(v1 - v0) * __fdividef(1.f, vs1 - vs0) : 0.f;  // table index per NDC unit
  const float vmax = fmaxf(fabsf(vs0), fabsf(vs1));
  const float v0f = (float)v0;
  float iu[3], rik[3], ge[3], av[3];
  bool up[3];                                        // the accepted side of edge k is "index >= root"
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    iu[k] = cols ? inv[k] : inv[3 + k];
    const float ik = cols ? inv[3 + k] : inv[k];
    rik[k] = __fdividef(1.f, ik);
    ge[k] = 1.7881393e-7f * fabsf(rik[k] * idx_per_v);   // 3 * 2^-24 / |i|, in index units
    av[k] = fabsf(ik) * vmax;
    up[k] = (ik > 0.f) == (idx_per_v >= 0.f);
  }
  for (int u = u0 + lane; u <= u1; u += 32) {
    const float un = us[u];
    float lo_f = v0f, hi_f = (float)v1;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float pk = un * iu[k], ck = inv[6 + k];
      const float root = fmaf(-(pk + ck) * rik[k] - vs0, idx_per_v, v0f);
      const float m = fmaf(fabsf(root), 3.8146973e-6f, fmaf((fabsf(pk) + fabsf(ck)) + av[k], ge[k], 0.05f));
      if (up[k]) lo_f = fmaxf(lo_f, root - m); else hi_f = fminf(hi_f, root + m);
    }
    if (!(lo_f <= hi_f)) continue;
    const int lo = (int)ceilf(lo_f), hi = (int)floorf(hi_f);
    for (int v = lo; v <= hi; ++v) {
      const int x = cols ? u : v, y = cols ? v : u;
      const float xn = xs[x], yn = ys[y];
      const float c0 = (xn * inv[0] + yn * inv[3]) + inv[6];
      const float c1 = (xn * inv[1] + yn * inv[4]) + inv[7];
      const float c2 = (xn * inv[2] + yn * inv[5]) + inv[8];
      if (c0 >= 0.f && c1 >= 0.f && c2 >= 0.f) {
        const float z = (c0 * zc[0] + c1 * zc[1]) + c2 * zc[2];
        const float zw = z * vp22 + vp23;
        put_key<K32>(keys_saddr, x * key_stride + y, zw, tri);
      }
    }
  }
}

struct V2Big {  // 64 bytes
  float inv[9];
  float zc[3];
  int tri;
  short x0, x1, y0, y1;
  int pad;
};

template <bool DEPTH, bool K32, int THREADS>
__global__ void __launch_bounds__(THREADS, K32 ? V2_K32_CTAS : JR_K64_CTAS)
k_vis2(const __grid_constant__ JrRenderArgs a, int tile_w, int tile_h, int tiles_x, int tiles_y) {
  static_assert(DEPTH || !K32, "z-only keys are for the depth shader");
  extern __shared__ __align__(16) unsigned char smem[];
  const V2Layout L = v2_layout(tile_w, tile_h, K32 ? 4 : 8);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem + L.keys);
  uint32_t* keys32 = reinterpret_cast<uint32_t*>(smem + L.keys);
  float* xs = reinterpret_cast<float*>(smem + L.xs);
  float* ys = reinterpret_cast<float*>(smem + L.ys);
  V2Big* bigq = reinterpret_cast<V2Big*>(smem + L.bigq);
  __shared__ int bigq_n;
  __shared__ int next_chunk;  // dynamic work distribution: next group of 32 triangles
  __shared__ int tri0_flag;
  __shared__ TriSetup tri0;
  __shared__ float s_w2c[16];
  __shared__ float s_vp[16];

  const int tid = threadIdx.x, lane = tid & 31;
  const int tiles = tiles_x * tiles_y;
  const int b = blockIdx.x / tiles;
  const int tile = blockIdx.x - b * tiles;
  const int tx0 = (tile / tiles_y) * tile_w;
  const int ty0 = (tile % tiles_y) * tile_h;
  const int tw = min(tile_w, a.W - tx0);
  const int th = min(tile_h, a.H - ty0);
  const uint32_t keys_saddr = (uint32_t)__cvta_generic_to_shared(keys);

  if (tid < 16) {
    s_w2c[tid] = a.world_to_clip.ptr[(long long)b * a.world_to_clip.batch_stride + tid];
    s_vp[tid] = a.viewport.ptr[(long long)b * a.viewport.batch_stride + tid];
  }
  if (tid == 0) { bigq_n = 0; tri0_flag = 0; next_chunk = 0; }
  {
    // all-ones = empty, whatever the key width (the layout rounds the buffer up to 16 bytes)
    const int n16 = (int)(L.xs >> 4);
    uint4* k4 = reinterpret_cast<uint4*>(smem + L.keys);
    for (int i = tid; i < n16; i += THREADS) k4[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
  }
  __syncthreads();
  for (int i = tid; i < tw; i += THREADS) xs[i] = ((float)(tx0 + i) - s_vp[3]) / s_vp[0];
  for (int i = tid; i < th; i += THREADS) ys[i] = ((float)(ty0 + i) - s_vp[7]) / s_vp[5];
  __syncthreads();

  const float vp00 = s_vp[0], vp03 = s_vp[3], vp11 = s_vp[5], vp13 = s_vp[7];
  const float vp22 = s_vp[10], vp23 = s_vp[11];
  const float* __restrict__ pos = a.position.ptr + (long long)b * a.position.batch_stride;
  const int32_t* __restrict__ faces = a.faces.ptr + (long long)b * a.faces.batch_stride;
  const float fx_lo = (float)tx0, fx_hi = (float)(tx0 + tw - 1);
  const float fy_lo = (float)ty0, fy_hi = (float)(ty0 + th - 1);

  // rasterise one small triangle with this lane
  auto raster_small = [&](const float* inv, const float* zc, int tri, int x0, int x1, int y0, int y1) {
    for (int x = x0; x <= x1; ++x) {
      const float xn = xs[x];
      const float pk0 = xn * inv[0], pk1 = xn * inv[1], pk2 = xn * inv[2];
      for (int y = y0; y <= y1; ++y) {
        const float yn = ys[y];
        const float c0 = (pk0 + yn * inv[3]) + inv[6];
        const float c1 = (pk1 + yn * inv[4]) + inv[7];
        const float c2 = (pk2 + yn * inv[5]) + inv[8];
        if (c0 >= 0.f && c1 >= 0.f && c2 >= 0.f) {
          const float z = (c0 * zc[0] + c1 * zc[1]) + c2 * zc[2];
          const float zw = z * vp22 + vp23;
          put_key<K32>(keys_saddr, x * tile_h + y, zw, (unsigned)tri);
        }
      }
    }
  };
  // flat raster of one bbox by a group of `nlanes` lanes (lane index `li`): pixel i -> (i / bh, i % bh)
  // with the quotient taken in fp32 ((i + 0.5) / bh is never within rounding distance of an integer
  // for i < 2^16, bh <= 255)
  auto raster_flat = [&](const float* inv, const float* zc, int tri, int x0, int y0, int bw, int bh, int li,
                         int nlanes) {
    const int n = bw * bh;
    const float rbh = 1.0f / (float)bh;
    for (int i = li; i < n; i += nlanes) {
      const int dx = (int)(((float)i + 0.5f) * rbh);
      const int x = x0 + dx, y = y0 + (i - dx * bh);
      const float xn = xs[x], yn = ys[y];
      const float c0 = (xn * inv[0] + yn * inv[3]) + inv[6];
      const float c1 = (xn * inv[1] + yn * inv[4]) + inv[7];
      const float c2 = (xn * inv[2] + yn * inv[5]) + inv[8];
      if (c0 >= 0.f && c1 >= 0.f && c2 >= 0.f) {
        const float z = (c0 * zc[0] + c1 * zc[1]) + c2 * zc[2];
        const float zw = z * vp22 + vp23;
        put_key<K32>(keys_saddr, x * tile_h + y, zw, (unsigned)tri);
      }
    }
  };

  // exact setup of one surviving triangle + dispatch by bbox size; called with warp-uniform control
  // flow (`valid` masks lanes without a record)
  auto fire = [&](bool valid, const float* M, const float* zc, int tri, unsigned bb) {
    float inv[9];
    const int x0 = bb & 0xff, x1 = (bb >> 8) & 0xff, y0 = (bb >> 16) & 0xff, y1 = bb >> 24;
    const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
    const int area = valid ? bw * bh : 0;
    if (valid) lu_inverse3(M, inv);
    bool is_medium = area > V2_SMALL_AREA && area <= V2_MEDIUM_AREA;
    if (area > V2_MEDIUM_AREA) {
      const int slot = atomicAdd(&bigq_n, 1);
      if (slot < V2_BIGCAP) {
        V2Big& q = bigq[slot];
#pragma unroll
        for (int k = 0; k < 9; ++k) q.inv[k] = inv[k];
        q.zc[0] = zc[0]; q.zc[1] = zc[1]; q.zc[2] = zc[2];
        q.tri = tri;
        q.x0 = (short)x0; q.x1 = (short)x1; q.y0 = (short)y0; q.y1 = (short)y1;
      } else {
        is_medium = true;  // queue full: the warp takes it
      }
    }
    if (area > 0 && area <= V2_SMALL_AREA) raster_small(inv, zc, tri, x0, x1, y0, y1);
    // medium triangles: one at a time, the whole warp on each
    unsigned mm = __ballot_sync(0xffffffffu, is_medium);
    while (mm) {
      const int src = __ffs(mm) - 1;
      mm &= mm - 1;
      float binv[9], bzc[3];
#pragma unroll
      for (int k = 0; k < 9; ++k) binv[k] = __shfl_sync(0xffffffffu, inv[k], src);
#pragma unroll
      for (int k = 0; k < 3; ++k) bzc[k] = __shfl_sync(0xffffffffu, zc[k], src);
      const int btri = __shfl_sync(0xffffffffu, tri, src);
      const unsigned bbb = __shfl_sync(0xffffffffu, bb, src);
      const int sx0 = bbb & 0xff, sx1 = (bbb >> 8) & 0xff, sy0 = (bbb >> 16) & 0xff, sy1 = bbb >> 24;
      if ((sx1 - sx0 + 1) * (sy1 - sy0 + 1) >= V2_HIER_AREA)
        raster_hier_warp<K32>(binv, bzc, (unsigned)btri, sx0, sy0, sx1, sy1, lane, xs, ys, keys_saddr, tile_h, vp22, vp23);
      else
        raster_flat(binv, bzc, btri, sx0, sy0, sx1 - sx0 + 1, sy1 - sy0 + 1, lane, 32);
    }
  };

  // Warps claim groups of 32 triangles dynamically (a static stride left warps that drew the
  // medium / ground triangles far behind the others at the barrier below).
  for (;;) {
    int t0 = 0;
    if (lane == 0) t0 = atomicAdd(&next_chunk, 32);
    t0 = __shfl_sync(0xffffffffu, t0, 0);
    if (t0 >= a.T) break;
    const int t = t0 + lane;
    bool surv = false;
    float M[9], zc[3];
    unsigned bb = 0;
    if (t < a.T) {
      // out-of-range indices are clamped (the reference's gathers clamp too) -- never read out of bounds
      const int vmax = a.n_pos - 1;
      const int i0 = min(max(faces[3 * t + 0], 0), vmax), i1 = min(max(faces[3 * t + 1], 0), vmax),
                i2 = min(max(faces[3 * t + 2], 0), vmax);
      const float p0x = pos[3 * i0], p0y = pos[3 * i0 + 1], p0z = pos[3 * i0 + 2];
      const float p1x = pos[3 * i1], p1y = pos[3 * i1 + 1], p1z = pos[3 * i1 + 2];
      const float p2x = pos[3 * i2], p2y = pos[3 * i2 + 1], p2z = pos[3 * i2 + 2];
      // rows 0, 1, 3 of to_clip (x, y, w); row 2 (z) only for survivors
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int rr = (r == 2) ? 3 : r;
        const float m0 = s_w2c[4 * rr], m1 = s_w2c[4 * rr + 1], m2 = s_w2c[4 * rr + 2], m3 = s_w2c[4 * rr + 3];
        M[0 + r] = ((p0x * m0 + p0y * m1) + p0z * m2) + m3;
        M[3 + r] = ((p1x * m0 + p1y * m1) + p1z * m2) + m3;
        M[6 + r] = ((p2x * m0 + p2y * m1) + p2z * m2) + m3;
      }
      const float det = det3(M);
      const bool cand = det > 1e-6f;  // keep & front (pipeline.py:98-100, :232)
      const bool fallback0 = DEPTH && (t == 0) && (det < -1e-6f);
      const float w0 = M[2], w1 = M[5], w2 = M[8];
      const bool behind = (w0 <= 0.f && w1 <= 0.f && w2 <= 0.f);
      if ((cand || fallback0) && !behind) {
        surv = true;
        int x0 = 0, x1 = tw - 1, y0 = 0, y1 = th - 1;
        if (w0 > 0.f && w1 > 0.f && w2 > 0.f && !fallback0) {
          const float r0 = __fdividef(1.f, w0), r1 = __fdividef(1.f, w1), r2 = __fdividef(1.f, w2);
          const float sx0 = (M[0] * r0) * vp00 + vp03, sx1 = (M[3] * r1) * vp00 + vp03, sx2 = (M[6] * r2) * vp00 + vp03;
          const float sy0 = (M[1] * r0) * vp11 + vp13, sy1 = (M[4] * r1) * vp11 + vp13, sy2 = (M[7] * r2) * vp11 + vp13;
          // conservative margin (see bbox_margin); fmaxf/fminf drop NaN towards "whole tile"
          const float mg = bbox_margin(sx0, sy0, sx1, sy1, sx2, sy2, 2.f * fmaxf(vp00, vp11));
          const float mnx = fmaxf(fminf(fminf(sx0, sx1), sx2) - mg, fx_lo);
          const float mxx = fminf(fmaxf(fmaxf(sx0, sx1), sx2) + mg, fx_hi);
          const float mny = fmaxf(fminf(fminf(sy0, sy1), sy2) - mg, fy_lo);
          const float mxy = fminf(fmaxf(fmaxf(sy0, sy1), sy2) + mg, fy_hi);
          if (!(mnx <= mxx) || !(mny <= mxy)) surv = false;
          x0 = (int)ceilf(mnx) - tx0; x1 = (int)floorf(mxx) - tx0;
          y0 = (int)ceilf(mny) - ty0; y1 = (int)floorf(mxy) - ty0;
          if (x0 > x1 || y0 > y1) surv = false;
        }
        if (surv) {
          const float m0 = s_w2c[8], m1 = s_w2c[9], m2 = s_w2c[10], m3 = s_w2c[11];
          zc[0] = ((p0x * m0 + p0y * m1) + p0z * m2) + m3;
          zc[1] = ((p1x * m0 + p1y * m1) + p1z * m2) + m3;
          zc[2] = ((p2x * m0 + p2y * m1) + p2z * m2) + m3;
          bb = (unsigned)x0 | ((unsigned)x1 << 8) | ((unsigned)y0 << 16) | ((unsigned)y1 << 24);
          if (fallback0) {
            // DepthShader quirk: kept back-facing triangle 0 fills pixels no candidate covers
            float inv[9];
            lu_inverse3(M, inv);
#pragma unroll
            for (int k = 0; k < 9; ++k) tri0.inv[k] = inv[k];
            tri0.zc[0] = zc[0]; tri0.zc[1] = zc[1]; tri0.zc[2] = zc[2];
            tri0.det = det;
            tri0_flag = 1;
            surv = false;
          }
        }
      }
    }
    // (Compacting survivors into full warps before `fire` was tried and measured SLOWER: +14 KB
    // shared memory per CTA shrinks L1 for the vertex gather and the per-lane raster loop still
    // runs to the longest box; see profiles/README.md.)
    if (__ballot_sync(0xffffffffu, surv)) fire(surv, M, zc, t, bb);
  }
  __syncthreads();
  const bool use0 = DEPTH && tri0_flag;
  const int npix_img = a.W * a.H;
  constexpr int PX = K32 ? 4 : 2;  // pixels per thread of the vectorised resolve
  // Single tile (tile-local index == pixel index), no triangle-0 fallback, column height a multiple of
  // PX (a thread's PX pixels share their x): the large triangles are folded into the resolve loop
  // below, against keys held in registers.
#ifdef JR_NO_FUSED_BIG
  const bool fused = false;
#else
  const bool fused = tiles == 1 && !use0 && (a.H % PX) == 0;
#endif
  // ---------------- large triangles: whole CTA, one at a time (a barrier between two entries makes
  // every pixel single-writer, so plain read-modify-write instead of a CAS loop)
  if (!fused) {
    const int nbig = min(bigq_n, V2_BIGCAP);
    for (int e = 0; e < nbig; ++e) {
      const V2Big& q = bigq[e];
      const float i0 = q.inv[0], i1 = q.inv[1], i2 = q.inv[2], i3 = q.inv[3], i4 = q.inv[4], i5 = q.inv[5],
                  i6 = q.inv[6], i7 = q.inv[7], i8 = q.inv[8];
      const float z0 = q.zc[0], z1 = q.zc[1], z2 = q.zc[2];
      const unsigned tri = (unsigned)q.tri;
      const int qx0 = q.x0, qy0 = q.y0, bh = q.y1 - q.y0 + 1;
      const int n = (q.x1 - q.x0 + 1) * bh;
      const float rbh = 1.0f / (float)bh;
      for (int i = tid; i < n; i += THREADS) {
        const int dx = (int)(((float)i + 0.5f) * rbh);
        const int x = qx0 + dx, y = qy0 + (i - dx * bh);
        const float xn = xs[x], yn = ys[y];
        const float c0 = (xn * i0 + yn * i3) + i6;
        const float c1 = (xn * i1 + yn * i4) + i7;
        const float c2 = (xn * i2 + yn * i5) + i8;
        if (c0 >= 0.f && c1 >= 0.f && c2 >= 0.f) {
          const float z = (c0 * z0 + c1 * z1) + c2 * z2;
          const float zw = z * vp22 + vp23;
          if (K32) {
            const uint32_t key = min(orderable(zw), 0xFFFFFFFEu);
            uint32_t* slot = &keys32[x * tile_h + y];
            if (key < *slot) *slot = key;
          } else {
            const unsigned long long key = ((unsigned long long)orderable(zw) << 32) | tri;
            unsigned long long* slot = &keys[x * tile_h + y];
            if (key < *slot) *slot = key;
          }
        }
      }
      __syncthreads();
    }
  }
  // ---------------- resolve
  int32_t* __restrict__ tri_out = a.tri_id ? a.tri_id + (long long)b * a.W * a.H : nullptr;
  float* __restrict__ z_out = DEPTH ? a.zbuffer + (long long)b * a.W * a.H : nullptr;
  if (fused) {
    // PX consecutive pixels per thread (128-bit LDS of their keys).  Each queued large triangle (the
    // ground plane of a Brax scene) is evaluated here, once per pixel, with the keys in registers:
    // no shared-memory read-modify-write, no barrier per triangle, the pixel coordinates are computed
    // once for all of them.  Same expressions, hence the same bits, as the rasterisers above.
    const int nbig = min(bigq_n, V2_BIGCAP);
    const int H = a.H;
    const float rH = 1.0f / (float)H;
    typedef typename std::conditional<K32, uint32_t, unsigned long long>::type KeyT;
    for (int i = tid; i < npix_img / PX; i += THREADS) {
      KeyT k[PX];
      if (K32) {
        const uint4 kk = reinterpret_cast<const uint4*>(keys32)[i];
        k[0] = kk.x; k[1] = kk.y; k[2 % PX] = kk.z; k[3 % PX] = kk.w;
      } else {
        const ulonglong2 kk = reinterpret_cast<const ulonglong2*>(keys)[i];
        k[0] = (KeyT)kk.x; k[1] = (KeyT)kk.y;
      }
      if (nbig) {
        const int p0 = i * PX;
        const int x = (int)(((float)p0 + 0.5f) * rH);  // exact for p0 < 2^16 (see raster_flat)
        const int y = p0 - x * H;                       // multiple of PX: the PX pixels are (x, y..y+PX-1)
        const float xn = xs[x];
        float yn[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) yn[p] = ys[y + p];
        for (int e = 0; e < nbig; ++e) {
          const V2Big& q = bigq[e];
          if (x < q.x0 || x > q.x1) continue;
          const float i3 = q.inv[3], i4 = q.inv[4], i5 = q.inv[5], i6 = q.inv[6], i7 = q.inv[7], i8 = q.inv[8];
          const float px0 = xn * q.inv[0], px1 = xn * q.inv[1], px2 = xn * q.inv[2];
          const float z0 = q.zc[0], z1 = q.zc[1], z2 = q.zc[2];
          const unsigned tri = (unsigned)q.tri;
          const int qy0 = q.y0;
          const unsigned qdy = (unsigned)(q.y1 - qy0);
#pragma unroll
          for (int p = 0; p < PX; ++p) {
            if ((unsigned)(y + p - qy0) > qdy) continue;
            const float c0 = (px0 + yn[p] * i3) + i6;
            const float c1 = (px1 + yn[p] * i4) + i7;
            const float c2 = (px2 + yn[p] * i5) + i8;
            if (c0 >= 0.f && c1 >= 0.f && c2 >= 0.f) {
              const float z = (c0 * z0 + c1 * z1) + c2 * z2;
              const float zw = z * vp22 + vp23;
              const KeyT key = K32 ? (KeyT)min(orderable(zw), 0xFFFFFFFEu)
                                   : (KeyT)(((unsigned long long)orderable(zw) << 32) | tri);
              if (key < k[p]) k[p] = key;
            }
          }
        }
      }
      if (K32) {
        const uint32_t k0 = (uint32_t)k[0], k1 = (uint32_t)k[1], k2 = (uint32_t)k[2 % PX], k3 = (uint32_t)k[3 % PX];
        if (k0 != ~0u && k1 != ~0u && k2 != ~0u && k3 != ~0u) {
          reinterpret_cast<float4*>(z_out)[i] =
              make_float4(from_orderable(k0), from_orderable(k1), from_orderable(k2), from_orderable(k3));
        } else {
          if (k0 != ~0u) z_out[4 * i] = from_orderable(k0);
          if (k1 != ~0u) z_out[4 * i + 1] = from_orderable(k1);
          if (k2 != ~0u) z_out[4 * i + 2] = from_orderable(k2);
          if (k3 != ~0u) z_out[4 * i + 3] = from_orderable(k3);
        }
      } else {
        const unsigned long long q0 = k[0], q1 = k[1];
        const bool e0 = q0 == ~0ull, e1 = q1 == ~0ull;
        if (DEPTH) {
          if (!e0 && !e1) {
            reinterpret_cast<float2*>(z_out)[i] =
                make_float2(from_orderable((uint32_t)(q0 >> 32)), from_orderable((uint32_t)(q1 >> 32)));
          } else {
            if (!e0) z_out[2 * i] = from_orderable((uint32_t)(q0 >> 32));
            if (!e1) z_out[2 * i + 1] = from_orderable((uint32_t)(q1 >> 32));
          }
        }
        if (tri_out)
          reinterpret_cast<int2*>(tri_out)[i] = make_int2(e0 ? -1 : (int)(unsigned)q0, e1 ? -1 : (int)(unsigned)q1);
      }
    }
  } else {
    const int dq = THREADS / th, dr = THREADS - dq * th;
    int lx = tid / th, ly = tid - lx * th;
    for (; lx < tw; lx += dq, ly += dr) {
      if (ly >= th) { ly -= th; ++lx; if (lx >= tw) break; }
      const long long pix = (long long)(tx0 + lx) * a.H + (ty0 + ly);
      int tri = -1;
      bool covered;
      if (K32) {
        const uint32_t key = keys32[lx * tile_h + ly];
        covered = key != ~0u;
        if (covered) z_out[pix] = from_orderable(key);
      } else {
        const unsigned long long key = keys[lx * tile_h + ly];
        covered = key != ~0ull;
        if (covered) {
          tri = (int)(unsigned)(key & 0xFFFFFFFFull);
          if (DEPTH) z_out[pix] = from_orderable((uint32_t)(key >> 32));
        }
      }
      if (covered) {
      } else if (use0) {
        float c[3];
        clip_coef(tri0.inv, xs[lx], ys[ly], c);
        if (c[0] >= 0.f && c[1] >= 0.f && c[2] >= 0.f) {
          const float z = (c[0] * tri0.zc[0] + c[1] * tri0.zc[1]) + c[2] * tri0.zc[2];
          z_out[pix] = z * vp22 + vp23;
          tri = 0;
        }
      }
      if (tri_out) tri_out[pix] = tri;
    }
  }
}

}  // namespace jr
