// jr_tiled.cuh -- two-level visibility for canvases that do not fit one shared-memory tile.
//
//   k_setup_bin      one thread per triangle (all images): vertex transform, cull, approximate
//                    bbox, exact LU inverse -> 64-byte record in the workspace; each warp then
//                    publishes, for every 64x64 screen tile its triangles touch, ONE 32-bit word
//                    (ballot of "lane's triangle overlaps the tile") into the tile's bitmask.
//                    Each (tile, word) has exactly one writer: no atomics, deterministic.
//   k_raster_tile    one CTA per (image, tile): keys of the tile in shared memory; warps scan the
//                    tile's bitmask words, compact the set bits into full warps of triangle ids,
//                    fetch the records and rasterise (small: lane; medium: warp; large: CTA),
//                    then resolve to tri_id (+ z for the depth shader).
//
// Triangle setup therefore runs once per triangle instead of once per (triangle, tile).
#pragma once
#include "jr_device.cuh"
#include "jr_geometry.cuh"
#include "jr_visibility.cuh"

namespace jr {

#ifndef JR_TL_TILE
#define JR_TL_TILE 64
#endif
constexpr int TL_TILE = JR_TL_TILE;  // power of two
constexpr int TL_SHIFT = (TL_TILE == 64) ? 6 : (TL_TILE == 32) ? 5 : (TL_TILE == 128) ? 7 : -1;
static_assert(TL_SHIFT > 0, "tile size must be 32, 64 or 128");
constexpr int TL_THREADS = 256;
constexpr int TL_BIGCAP = 32;
#ifndef JR_TL_BIG_AREA
#define JR_TL_BIG_AREA 1024
#endif
constexpr int TL_BIG_AREA = JR_TL_BIG_AREA;  // bbox (clipped to the tile) above this: the whole CTA sweeps it
#ifndef JR_TL_MASKCAP
#define JR_TL_MASKCAP 1024
#endif
#ifndef JR_TL_SPAN
#define JR_TL_SPAN 2      // warp-cooperative boxes of >= V2_HIER_AREA pixels: span raster from the analytic roots (1: exact
                          // interval search, 0: hierarchical block raster)
#endif
#ifndef JR_TL_SPAN_AREA
#define JR_TL_SPAN_AREA 256   // queued boxes from this many pixels: span / hierarchical raster; below: lane per pixel
#endif
#ifndef JR_TL_BIG_FILL
#define JR_TL_BIG_FILL 6  // of 16 samples: boxes above TL_BIG_AREA go to the CTA-wide sweep only when this full
#endif
#ifndef JR_TL_CTAS
#define JR_TL_CTAS 4
#endif
#ifndef JR_TL_K32_CTAS
#define JR_TL_K32_CTAS 5
#endif
constexpr int TL_MASKCAP = JR_TL_MASKCAP;  // bitmask words staged in shared memory per chunk (x32 triangles)

struct __align__(16) TriRecord {  // 64 bytes
  float inv[9];
  float zc[3];
  unsigned short x0, x1, y0, y1;  // absolute pixel bbox, inclusive
  int flags;                      // 1 = DepthShader triangle-0 fallback record
  int pad;
};
static_assert(sizeof(TriRecord) == 64, "TriRecord must be 64 bytes");

struct TiledLayout {
  int tiles_x, tiles_y, tiles, words;
  size_t rec, mask, total;
};
__host__ __device__ inline TiledLayout tiled_layout(int B, int W, int H, int T) {
  TiledLayout L;
  L.tiles_x = (W + TL_TILE - 1) / TL_TILE;
  L.tiles_y = (H + TL_TILE - 1) / TL_TILE;
  L.tiles = L.tiles_x * L.tiles_y;
  L.words = (T + 31) / 32;
  L.rec = 0;
  size_t r = (size_t)B * (size_t)(T > 0 ? T : 1) * sizeof(TriRecord);
  L.mask = (r + 255) & ~(size_t)255;
  L.total = L.mask + (size_t)B * L.tiles * (L.words > 0 ? L.words : 1) * 4;
  return L;
}

template <bool DEPTH>
__global__ void __launch_bounds__(256)
k_setup_bin(const __grid_constant__ JrRenderArgs a, TriRecord* __restrict__ recs, unsigned* __restrict__ masks,
            TiledLayout L) {
  __shared__ float s_w2c[16];
  __shared__ float s_vp[16];
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < 16) {
    s_w2c[tid] = a.world_to_clip.ptr[(long long)b * a.world_to_clip.batch_stride + tid];
    s_vp[tid] = a.viewport.ptr[(long long)b * a.viewport.batch_stride + tid];
  }
  __syncthreads();
  const float vp00 = s_vp[0], vp03 = s_vp[3], vp11 = s_vp[5], vp13 = s_vp[7];
  const float* __restrict__ pos = a.position.ptr + (long long)b * a.position.batch_stride;
  const int32_t* __restrict__ faces = a.faces.ptr + (long long)b * a.faces.batch_stride;
  const int t = blockIdx.x * 256 + tid;
  bool surv = false;
  bool wrote = false;  // this thread stored its record
  int x0 = 0, x1 = a.W - 1, y0 = 0, y1 = a.H - 1;
  if (t < a.T) {
    const int vmax = a.n_pos - 1;  // out-of-range indices are clamped, never read out of bounds
    const int i0 = min(max(faces[3 * t + 0], 0), vmax), i1 = min(max(faces[3 * t + 1], 0), vmax),
              i2 = min(max(faces[3 * t + 2], 0), vmax);
    const Vec3 q0 = fetch_position(a, b, pos, i0), q1 = fetch_position(a, b, pos, i1), q2 = fetch_position(a, b, pos, i2);
    const float p0x = q0.x, p0y = q0.y, p0z = q0.z;
    const float p1x = q1.x, p1y = q1.y, p1z = q1.z;
    const float p2x = q2.x, p2y = q2.y, p2z = q2.z;
    float M[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int rr = (r == 2) ? 3 : r;
      const float m0 = s_w2c[4 * rr], m1 = s_w2c[4 * rr + 1], m2 = s_w2c[4 * rr + 2], m3 = s_w2c[4 * rr + 3];
      M[0 + r] = ((p0x * m0 + p0y * m1) + p0z * m2) + m3;
      M[3 + r] = ((p1x * m0 + p1y * m1) + p1z * m2) + m3;
      M[6 + r] = ((p2x * m0 + p2y * m1) + p2z * m2) + m3;
    }
    const float det = det3(M);
    const bool cand = det > 1e-6f;
    const bool fallback0 = DEPTH && (t == 0) && (det < -1e-6f);
    const float w0 = M[2], w1 = M[5], w2 = M[8];
    const bool behind = (w0 <= 0.f && w1 <= 0.f && w2 <= 0.f);
    if ((cand || fallback0) && !behind) {
      surv = true;
      if (w0 > 0.f && w1 > 0.f && w2 > 0.f && !fallback0) {
        const float r0 = __fdividef(1.f, w0), r1 = __fdividef(1.f, w1), r2 = __fdividef(1.f, w2);
        const float sx0 = (M[0] * r0) * vp00 + vp03, sx1 = (M[3] * r1) * vp00 + vp03, sx2 = (M[6] * r2) * vp00 + vp03;
        const float sy0 = (M[1] * r0) * vp11 + vp13, sy1 = (M[4] * r1) * vp11 + vp13, sy2 = (M[7] * r2) * vp11 + vp13;
        const float mg = bbox_margin(sx0, sy0, sx1, sy1, sx2, sy2, 2.f * fmaxf(vp00, vp11));
        const float mnx = fmaxf(fminf(fminf(sx0, sx1), sx2) - mg, 0.f);
        const float mxx = fminf(fmaxf(fmaxf(sx0, sx1), sx2) + mg, (float)(a.W - 1));
        const float mny = fmaxf(fminf(fminf(sy0, sy1), sy2) - mg, 0.f);
        const float mxy = fminf(fmaxf(fmaxf(sy0, sy1), sy2) + mg, (float)(a.H - 1));
        if (!(mnx <= mxx) || !(mny <= mxy)) surv = false;
        x0 = (int)ceilf(mnx); x1 = (int)floorf(mxx);
        y0 = (int)ceilf(mny); y1 = (int)floorf(mxy);
        if (x0 > x1 || y0 > y1) surv = false;
      }
      if (surv) {
        TriRecord r;
        lu_inverse3(M, r.inv);
        const float m0 = s_w2c[8], m1 = s_w2c[9], m2 = s_w2c[10], m3 = s_w2c[11];
        r.zc[0] = ((p0x * m0 + p0y * m1) + p0z * m2) + m3;
        r.zc[1] = ((p1x * m0 + p1y * m1) + p1z * m2) + m3;
        r.zc[2] = ((p2x * m0 + p2y * m1) + p2z * m2) + m3;
        r.x0 = (unsigned short)x0; r.x1 = (unsigned short)x1; r.y0 = (unsigned short)y0; r.y1 = (unsigned short)y1;
        r.flags = fallback0 ? 1 : 0;
        r.pad = 0;
        float4* dst = reinterpret_cast<float4*>(recs + (size_t)b * a.T + t);
        const float4* src = reinterpret_cast<const float4*>(&r);
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
        wrote = true;
        if (fallback0) surv = false;  // not a candidate: never enters a bin
      }
    }
    // record 0's flag is always defined (the raster kernel looks at it for the DepthShader quirk), also when
    // triangle 0 is front-facing but its bbox was culled: the workspace is never cleared, a stale flag from an
    // earlier call would switch the fallback on
    if (DEPTH && t == 0 && !wrote) recs[(size_t)b * a.T].flags = 0;
  }
  // ---- publish tile bitmask words (one writer per (tile, word))
  int tx0 = x0 / TL_TILE, tx1 = x1 / TL_TILE, ty0 = y0 / TL_TILE, ty1 = y1 / TL_TILE;
  if (!surv) { tx0 = 1 << 20; tx1 = -1; ty0 = 1 << 20; ty1 = -1; }
  int ux0 = tx0, ux1 = tx1, uy0 = ty0, uy1 = ty1;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    ux0 = min(ux0, __shfl_xor_sync(0xffffffffu, ux0, off));
    ux1 = max(ux1, __shfl_xor_sync(0xffffffffu, ux1, off));
    uy0 = min(uy0, __shfl_xor_sync(0xffffffffu, uy0, off));
    uy1 = max(uy1, __shfl_xor_sync(0xffffffffu, uy1, off));
  }
  if (ux1 < ux0) return;  // no survivor in this warp
  const int word = t >> 5;
  unsigned* mrow = masks + (size_t)b * L.tiles * L.words;
  for (int tx = ux0; tx <= ux1; ++tx)
    for (int ty = uy0; ty <= uy1; ++ty) {
      const bool hit = surv && tx >= tx0 && tx <= tx1 && ty >= ty0 && ty <= ty1;
      const unsigned w = __ballot_sync(0xffffffffu, hit);
      if (w && lane == 0) mrow[(size_t)(tx * L.tiles_y + ty) * L.words + word] = w;
    }
}

#ifndef JR_TL_MEDCAP
#define JR_TL_MEDCAP 2048
#endif
constexpr int TL_MEDCAP = JR_TL_MEDCAP;  // warp-cooperative triangles queued per tile for phase 2 (overflow: rasterised in place)
struct TLSmem { size_t keys, xs, ys, ring, bigq, mask, nz, medq, total; };
__host__ __device__ inline TLSmem tl_smem(int key_bytes) {
  TLSmem S;
  S.keys = 0;
  S.xs = (size_t)TL_TILE * TL_TILE * key_bytes;
  S.ys = S.xs + TL_TILE * 4;
  S.ring = S.ys + TL_TILE * 4;
  S.bigq = S.ring + (TL_THREADS / 32) * 64 * 4;
  S.mask = S.bigq + (size_t)TL_BIGCAP * 64;
  S.nz = S.mask + (size_t)TL_MASKCAP * 4;
  S.medq = S.nz + (size_t)TL_MASKCAP * 2;
  S.total = S.medq + (size_t)TL_MEDCAP * 4;
  return S;
}

// Visible-triangle lists of the attribute stage, built by the resolve (non-depth passes; all NULL: not wanted).  The
// first pixel of a run of equal ids in a column flags its triangle (atomicOr on the image's flag bits, workspace) and
// the first to flag it appends it to the image's list -- k_mark_visible's logic without its re-read of the G-buffer.
struct TLVis {
  unsigned* flags;  // B*T bits, then B counters (zeroed by the caller before the launch)
  int* list;        // (B, T)
  int* count;       // (B)
  int* slot_map;    // (B, T) triangle -> record slot (compact records), or NULL
};

// K32: depth shader without a triangle-id output (shadow passes) -- z-only 32-bit keys, see k_vis2.
template <bool DEPTH, bool K32>
__global__ void __launch_bounds__(TL_THREADS, K32 ? JR_TL_K32_CTAS : JR_TL_CTAS)
k_raster_tile(const __grid_constant__ JrRenderArgs a, const TriRecord* __restrict__ recs,
              const unsigned* __restrict__ masks, TiledLayout L, TLVis vis) {
  static_assert(DEPTH || !K32, "z-only keys are for the depth shader");
  extern __shared__ __align__(16) unsigned char smem[];
  const TLSmem S = tl_smem(K32 ? 4 : 8);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem + S.keys);
  uint32_t* keys32 = reinterpret_cast<uint32_t*>(smem + S.keys);
  float* xs = reinterpret_cast<float*>(smem + S.xs);
  float* ys = reinterpret_cast<float*>(smem + S.ys);
  V2Big* bigq = reinterpret_cast<V2Big*>(smem + S.bigq);
  unsigned* s_mask = reinterpret_cast<unsigned*>(smem + S.mask);
  unsigned short* s_nz = reinterpret_cast<unsigned short*>(smem + S.nz);   // indices of the chunk's non-zero words
  int* medq = reinterpret_cast<int*>(smem + S.medq);                       // warp-cooperative triangles (phase 2)
  __shared__ int s_nzn, s_next, s_medn, s_mednext;
  __shared__ int bigq_n;
  __shared__ int tri0_flag;
  __shared__ TriSetup tri0;
  __shared__ float s_vp[16];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int* ring = reinterpret_cast<int*>(smem + S.ring) + warp * 64;
  const int b = blockIdx.x / L.tiles;
  const int tile = blockIdx.x - b * L.tiles;
  const int tx0 = (tile / L.tiles_y) * TL_TILE;
  const int ty0 = (tile % L.tiles_y) * TL_TILE;
  const int tw = min(TL_TILE, a.W - tx0);
  const int th = min(TL_TILE, a.H - ty0);
  const uint32_t keys_saddr = (uint32_t)__cvta_generic_to_shared(keys);
  const TriRecord* __restrict__ rec_b = recs + (size_t)b * a.T;

  if (tid < 16) s_vp[tid] = a.viewport.ptr[(long long)b * a.viewport.batch_stride + tid];
  if (tid == 0) {
    bigq_n = 0;
    tri0_flag = 0;
    s_medn = 0;
    s_mednext = 0;
  }
  {
    uint4* k4 = reinterpret_cast<uint4*>(smem + S.keys);
    for (int i = tid; i < (int)(S.xs >> 4); i += TL_THREADS) k4[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
  }
  __syncthreads();
  for (int i = tid; i < tw; i += TL_THREADS) xs[i] = ((float)(tx0 + i) - s_vp[3]) / s_vp[0];
  for (int i = tid; i < th; i += TL_THREADS) ys[i] = ((float)(ty0 + i) - s_vp[7]) / s_vp[5];
  __syncthreads();
  const float vp22 = s_vp[10], vp23 = s_vp[11];

  auto put = [&](int x, int y, const float* inv, const float* zc, unsigned tri) {
    const float xn = xs[x], yn = ys[y];
    const float c0 = (xn * inv[0] + yn * inv[3]) + inv[6];
    const float c1 = (xn * inv[1] + yn * inv[4]) + inv[7];
    const float c2 = (xn * inv[2] + yn * inv[5]) + inv[8];
    if (c0 >= 0.f && c1 >= 0.f && c2 >= 0.f) {
      const float z = (c0 * zc[0] + c1 * zc[1]) + c2 * zc[2];
      const float zw = z * vp22 + vp23;
      put_key<K32>(keys_saddr, x * TL_TILE + y, zw, tri);
    }
  };

  // one warp, one triangle (warp-uniform arguments, tile-local inclusive box of more than V2_SMALL_AREA pixels)
  auto warp_raster = [&](const float* binv, const float* bzc, unsigned btri, int sx0, int sy0, int sx1, int sy1) {
    const int sbh = sy1 - sy0 + 1;
    const int n = (sx1 - sx0 + 1) * sbh;
    if (n >= JR_TL_SPAN_AREA) {
      if (JR_TL_SPAN == 2)
        raster_span2_warp<K32>(binv, bzc, btri, sx0, sy0, sx1, sy1, lane, xs, ys, keys_saddr, TL_TILE, vp22, vp23);
      else if (JR_TL_SPAN == 1)
        raster_span_warp<K32>(binv, bzc, btri, sx0, sy0, sx1, sy1, lane, xs, ys, keys_saddr, TL_TILE, vp22, vp23);
      else
        raster_hier_warp<K32>(binv, bzc, btri, sx0, sy0, sx1, sy1, lane, xs, ys, keys_saddr, TL_TILE, vp22, vp23);
    } else {
      const float rbh = 1.0f / (float)sbh;
      for (int i = lane; i < n; i += 32) {
        const int dx = (int)(((float)i + 0.5f) * rbh);
        put(sx0 + dx, sy0 + (i - dx * sbh), binv, bzc, btri);
      }
    }
  };

  // fetch the record of triangle `t`, clip its bbox to the tile, dispatch by size (warp-uniform call)
  auto fire = [&](bool valid, int t) {
    float inv[9], zc[3];
    int x0 = 0, x1 = -1, y0 = 0, y1 = -1;
    if (valid) {
      const float4* src = reinterpret_cast<const float4*>(rec_b + t);
      const float4 q0 = src[0], q1 = src[1], q2 = src[2], q3 = src[3];
      inv[0] = q0.x; inv[1] = q0.y; inv[2] = q0.z; inv[3] = q0.w;
      inv[4] = q1.x; inv[5] = q1.y; inv[6] = q1.z; inv[7] = q1.w;
      inv[8] = q2.x; zc[0] = q2.y; zc[1] = q2.z; zc[2] = q2.w;
      const unsigned bx = __float_as_uint(q3.x), by = __float_as_uint(q3.y);
      x0 = max((int)(bx & 0xffff) - tx0, 0); x1 = min((int)(bx >> 16) - tx0, tw - 1);
      y0 = max((int)(by & 0xffff) - ty0, 0); y1 = min((int)(by >> 16) - ty0, th - 1);
    }
    const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
    const int area = (valid && bw > 0 && bh > 0) ? bw * bh : 0;
    // above the per-lane size the whole warp rasterises the triangle (hierarchically from V2_HIER_AREA
    // pixels); triangles covering a large part of the tile (the ground plane: most tiles of a Brax
    // frame hold nothing else) are queued for the whole CTA -- one warp sweeping 2 x 4096 pixels while
    // the other seven idle was the critical path of those tiles
    bool is_medium = area > V2_SMALL_AREA;
    bool is_big = area > TL_BIG_AREA;
    if (JR_TL_SPAN && is_big) {
      // the CTA-wide sweep visits every pixel of the box: it is for triangles that FILL theirs (the ground plane).
      // A thin triangle with a large box (fill ~7 %) is left to the warp's span raster.  4 x 4 samples decide.
      int inside = 0;
#pragma unroll 1
      for (int s = 0; s < 16; ++s) {
        const int sx = x0 + ((2 * (s >> 2) + 1) * (bw - 1)) / 8, sy = y0 + ((2 * (s & 3) + 1) * (bh - 1)) / 8;
        const float xn = xs[sx], yn = ys[sy];
        inside += (((xn * inv[0] + yn * inv[3]) + inv[6]) >= 0.f && ((xn * inv[1] + yn * inv[4]) + inv[7]) >= 0.f &&
                   ((xn * inv[2] + yn * inv[5]) + inv[8]) >= 0.f) ? 1 : 0;
      }
      is_big = inside >= JR_TL_BIG_FILL;
    }
    if (is_big) {
      const int slot = atomicAdd(&bigq_n, 1);
      if (slot < TL_BIGCAP) {
        V2Big& q = bigq[slot];
#pragma unroll
        for (int k = 0; k < 9; ++k) q.inv[k] = inv[k];
        q.zc[0] = zc[0]; q.zc[1] = zc[1]; q.zc[2] = zc[2];
        q.tri = t;
        q.x0 = (short)x0; q.x1 = (short)x1; q.y0 = (short)y0; q.y1 = (short)y1;
        is_medium = false;
      }
    }
    if (area > 0 && area <= V2_SMALL_AREA) {
      for (int x = x0; x <= x1; ++x)
        for (int y = y0; y <= y1; ++y) put(x, y, inv, zc, (unsigned)t);
    }
    const unsigned lt = (1u << lane) - 1u;
    unsigned mm = __ballot_sync(0xffffffffu, is_medium);
    if (mm) {
      // warp-cooperative boxes are not rasterised here: a warp whose 32 triangles hold a dozen of them kept the other
      // seven waiting at the barrier (half of all stall samples).  They are queued; phase 2 hands them out one by one.
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_medn, __popc(mm));
      base = __shfl_sync(0xffffffffu, base, 0);
      const int slot = base + __popc(mm & lt);
      if (is_medium && slot < TL_MEDCAP) { medq[slot] = t; is_medium = false; }
      mm = __ballot_sync(0xffffffffu, is_medium);   // queue full: in place, as before
    }
    while (mm) {
      const int src = __ffs(mm) - 1;
      mm &= mm - 1;
      float binv[9], bzc[3];
#pragma unroll
      for (int k = 0; k < 9; ++k) binv[k] = __shfl_sync(0xffffffffu, inv[k], src);
#pragma unroll
      for (int k = 0; k < 3; ++k) bzc[k] = __shfl_sync(0xffffffffu, zc[k], src);
      const unsigned btri = (unsigned)__shfl_sync(0xffffffffu, t, src);
      warp_raster(binv, bzc, btri, __shfl_sync(0xffffffffu, x0, src), __shfl_sync(0xffffffffu, y0, src),
                  __shfl_sync(0xffffffffu, x1, src), __shfl_sync(0xffffffffu, y1, src));
    }
  };

  // DepthShader triangle-0 fallback record
  if (DEPTH && tid == 0 && a.T > 0) {
    const TriRecord& r0 = rec_b[0];
    if (r0.flags == 1) {  // k_setup_bin always defines record 0's flag
#pragma unroll
      for (int k = 0; k < 9; ++k) tri0.inv[k] = r0.inv[k];
      tri0.zc[0] = r0.zc[0]; tri0.zc[1] = r0.zc[1]; tri0.zc[2] = r0.zc[2];
      tri0_flag = 1;
    }
  }

  // ---- scan the tile's bitmask; compact set bits into full warps of triangle ids
  const unsigned* __restrict__ mrow = masks + ((size_t)b * L.tiles + tile) * L.words;
  int head = 0, pc = 0;  // warp-uniform ring state
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int chunk0 = 0; chunk0 < L.words; chunk0 += TL_MASKCAP) {
    // stage the tile's bitmask words in shared memory (coalesced; a per-word global load in the
    // scan loop below serialised ~80 dependent DRAM round trips per warp)
    const int n = min(TL_MASKCAP, L.words - chunk0);
    if (tid == 0) { s_nzn = 0; s_next = 0; }
    __syncthreads();
    // stage the words and list the non-zero ones (most tiles of a frame see two triangles of 20 000: stepping through
    // 625 empty words per warp-eighth was 11-13 % of the kernel's instructions)
    for (int i0 = 0; i0 < n; i0 += TL_THREADS) {
      const int i = i0 + tid;
      const unsigned w = i < n ? mrow[chunk0 + i] : 0u;
      if (i < n) s_mask[i] = w;
      const unsigned nzb = __ballot_sync(0xffffffffu, w != 0u);
      if (nzb) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_nzn, __popc(nzb));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (w != 0u) s_nz[base + __popc(nzb & lt_mask)] = (unsigned short)i;
      }
    }
    __syncthreads();
    // warps claim non-zero words one at a time (the order is free: the key updates are minima)
    const int nzn = s_nzn;
    for (;;) {
      int e = 0;
      if (lane == 0) e = atomicAdd(&s_next, 1);
      e = __shfl_sync(0xffffffffu, e, 0);
      if (e >= nzn) break;
      const int i = s_nz[e];
      const unsigned w = s_mask[i];
      if ((w >> lane) & 1u) ring[(head + pc + __popc(w & lt_mask)) & 63] = (chunk0 + i) * 32 + lane;
      pc += __popc(w);
      __syncwarp();
      if (pc >= 32) {
        const int t = ring[(head + lane) & 63];
        head = (head + 32) & 63;
        pc -= 32;
        __syncwarp();
        fire(true, t);
      }
    }
    __syncthreads();
  }
  if (pc > 0) {
    const int t = (lane < pc) ? ring[(head + lane) & 63] : 0;
    fire(lane < pc, t);
  }
  __syncthreads();
  // ---- phase 2: the queued warp-cooperative triangles, one per claim (every lane fetches the same record: one
  // broadcast transaction per 16 bytes instead of 16 shuffles)
  {
    const int nmed = min(s_medn, TL_MEDCAP);
    for (;;) {
      int e = 0;
      if (lane == 0) e = atomicAdd(&s_mednext, 1);
      e = __shfl_sync(0xffffffffu, e, 0);
      if (e >= nmed) break;
      const int t = medq[e];
      const float4* src = reinterpret_cast<const float4*>(rec_b + t);
      const float4 q0 = src[0], q1 = src[1], q2 = src[2], q3 = src[3];
      const float inv[9] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x};
      const float zc[3] = {q2.y, q2.z, q2.w};
      const unsigned bx = __float_as_uint(q3.x), by = __float_as_uint(q3.y);
      warp_raster(inv, zc, (unsigned)t, max((int)(bx & 0xffff) - tx0, 0), max((int)(by & 0xffff) - ty0, 0),
                  min((int)(bx >> 16) - tx0, tw - 1), min((int)(by >> 16) - ty0, th - 1));
    }
    __syncthreads();
  }
  // ---- large triangles: whole CTA, one at a time, single writer per pixel
  {
    const int nbig = min(bigq_n, TL_BIGCAP);
    for (int e = 0; e < nbig; ++e) {
      const V2Big& q = bigq[e];
      const float i0 = q.inv[0], i1 = q.inv[1], i2 = q.inv[2], i3 = q.inv[3], i4 = q.inv[4], i5 = q.inv[5],
                  i6 = q.inv[6], i7 = q.inv[7], i8 = q.inv[8];
      const float z0 = q.zc[0], z1 = q.zc[1], z2 = q.zc[2];
      const unsigned tri = (unsigned)q.tri;
      const int qx0 = q.x0, qy0 = q.y0, bh = q.y1 - q.y0 + 1;
      const int n = (q.x1 - q.x0 + 1) * bh;
      const float rbh = 1.0f / (float)bh;
      for (int i = tid; i < n; i += TL_THREADS) {
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
            uint32_t* slot = &keys32[x * TL_TILE + y];
            if (key < *slot) *slot = key;
          } else {
            const unsigned long long key = ((unsigned long long)orderable(zw) << 32) | tri;
            unsigned long long* slot = &keys[x * TL_TILE + y];
            if (key < *slot) *slot = key;
          }
        }
      }
      __syncthreads();
    }
  }
  // ---- resolve
  int32_t* __restrict__ tri_out = a.tri_id ? a.tri_id + (long long)b * a.W * a.H : nullptr;
  float* __restrict__ z_out = DEPTH ? a.zbuffer + (long long)b * a.W * a.H : nullptr;
  const bool use0 = DEPTH && tri0_flag;
  for (int i = tid; i < tw * TL_TILE; i += TL_THREADS) {
    const int lx = i >> TL_SHIFT, ly = i & (TL_TILE - 1);
    if (ly >= th) continue;
    const long long pix = (long long)(tx0 + lx) * a.H + (ty0 + ly);
    int tri = -1;
    bool covered, wrote = false;
    float zv = a.depth_fill_value;
    if (K32) {
      const uint32_t key = keys32[i];
      covered = key != ~0u;
      if (covered) { zv = from_orderable(key); wrote = true; }
    } else {
      const unsigned long long key = keys[i];
      covered = key != ~0ull;
      if (covered) {
        tri = (int)(unsigned)(key & 0xFFFFFFFFull);
        if (DEPTH) { zv = from_orderable((uint32_t)(key >> 32)); wrote = true; }
        if (!DEPTH && vis.flags && !(ly > 0 && (unsigned)(keys[i - 1] & 0xFFFFFFFFull) == (unsigned)tri)) {
          // first pixel of a run of this id in the column (an empty predecessor reads 0xFFFFFFFF: never a triangle)
          const long long bit = (long long)b * a.T + tri;
          unsigned* w = vis.flags + (bit >> 5);
          const unsigned m = 1u << (bit & 31);
          if (!(*w & m) && !(atomicOr(w, m) & m)) {
            const int slot = atomicAdd(&vis.count[b], 1);
            vis.list[(long long)b * a.T + slot] = tri;
            if (vis.slot_map) vis.slot_map[(long long)b * a.T + tri] = slot;
          }
        }
      }
    }
    if (covered) {
    } else if (use0) {
      float c[3];
      clip_coef(tri0.inv, xs[lx], ys[ly], c);
      if (c[0] >= 0.f && c[1] >= 0.f && c[2] >= 0.f) {
        const float z = (c[0] * tri0.zc[0] + c[1] * tri0.zc[1]) + c[2] * tri0.zc[2];
        zv = z * vp22 + vp23; wrote = true;
        tri = 0;
      }
    }
    // depth epilogue (jr_b200.h): + offset on every written depth, optional fill of the uncovered pixels
    if (DEPTH && (wrote || a.depth_fill != 0)) z_out[pix] = a.depth_offset != 0.f ? zv + a.depth_offset : zv;
    if (tri_out) tri_out[pix] = tri;
  }
}

}  // namespace jr
