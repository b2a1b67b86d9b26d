// jr_vis3.cuh -- k_vis3: filtered two-phase visibility kernel for canvases that fit one shared-memory tile
// (the bench workload, every 84x84 / 32x32 Brax frame and their shadow passes).
//
// One CTA (256 threads) per image, the image's depth keys in shared memory, as in k_vis2.  What changed
// is WHERE the exact (no-FMA, IEEE-division, LAPACK-order) arithmetic of the reference runs
// (renderer/pipeline.py:76-113, :163-279):
//
//   phase A  FILTER, lane = triangle, all T triangles.  Vertices arrive through a coalesced stage when
//            a warp's 32 triangles read 96 consecutive vertices (corner-expanded meshes, what
//            Renderer.render produces: renderer.py:277-296), else by gather.  FMA transform of x / y / w,
//            FMA determinant, reciprocal-based screen bbox.  A triangle is DROPPED only when that is
//            certain under rounding-error bounds that cover both this evaluation and the reference's:
//               all w < -e_w                       (behind the camera: pipeline.py:232 w/ :98-100)
//               det + e_det < 1e-6                 (not "keep & front-facing")
//               all w > e_w and the bbox, grown by the conservative margin, holds no sample
//            everything else -- including every uncertain case -- is pushed, as a 16-bit triangle
//            offset, into one of four CTA-wide lists by bbox size.
//   phase C  EXACT, lane = surviving triangle, full warps.  Re-gathers the vertices and runs exactly the
//            arithmetic of k_vis2 (the trusted path): no-FMA clip transform, jax `_det_3x3` determinant,
//            cull, bbox, LAPACK-order LU inverse, edge functions, depth, key atomics.  Phase A never
//            decides anything that reaches a pixel; the results are bit-identical to k_vis2.
//   resolve  keys -> z (+ triangle id), large triangles folded in (k_vis2's resolve).
//
// At 84x84 a bench image (ground cube + 10 capsules) has 1932 triangles of which ~480 survive phase A (~930 are
// front-facing), ~410 the exact cull and ~150 cover a sample: the exact set-up (LU inverse: 9 IEEE divisions) runs in
// ~18 well-filled warp rounds per image instead of 61 half-empty ones, and the per-lane raster loops are grouped by
// box size.  Variants: INST (geometry instanced at the fetch), CLUSTER (one image split over a 2-CTA cluster, small
// batches), STATS (counters); non-depth builds also emit the visible-triangle lists (V3Vis).  What was measured and
// rejected (TMA / register prefetch of phase A, a per-sample coverage test in the filter, 5 CTAs/SM): profiles/README.md.
#pragma once
#include <type_traits>
#include "jr_device.cuh"
#include "jr_geometry.cuh"
#include "jr_visibility.cuh"

namespace jr {

constexpr int V3_THREADS = 256;
constexpr int V3_NW = V3_THREADS / 32;
#ifndef JR_V3_K32_CTAS
#define JR_V3_K32_CTAS 4
#endif
#ifndef JR_V3_GRAIN3
#define JR_V3_GRAIN3 8
#endif
#ifndef JR_V3_GRAIN2
#define JR_V3_GRAIN2 32
#endif
#ifndef JR_V3_RESOLVE_UNROLL
#define JR_V3_RESOLVE_UNROLL 4
#endif
#ifndef JR_V3_NV
#define JR_V3_NV 1
#endif
#ifndef JR_V3_K64_CTAS
#define JR_V3_K64_CTAS 4
#endif
// survivor lists by bbox size (pixels): <= 4, <= 16, <= 64, larger (incl. "whole tile")
constexpr int V3_A0 = 4, V3_A1 = 16, V3_A2 = 64;
constexpr int V3_CAP0 = 1024, V3_CAP1 = 512, V3_CAP2 = 512, V3_CAP3 = 512;
constexpr int V3_OFF0 = 0, V3_OFF1 = V3_CAP0, V3_OFF2 = V3_OFF1 + V3_CAP1, V3_OFF3 = V3_OFF2 + V3_CAP2;
constexpr int V3_LIST_ENTRIES = V3_OFF3 + V3_CAP3;
constexpr int V3_BIGCAP = 16;
constexpr int V3_STAGE_WORDS = 288;      // 32 triangles x 3 vertices x 3 floats
constexpr int V3_TMAX16 = 65535;         // 16-bit list entries: triangles beyond are set up in place

// Shared-memory layout: everything but the key tile sits at compile-time offsets (the compiler re-derives
// run-time offsets inside the hot loops when registers are short).  The span table of the large triangles
// (resolve) re-uses lists + stage, which are dead by then.
constexpr int V3_SM_BIGQ = 0;
constexpr int V3_SM_LISTS = V3_SM_BIGQ + V3_BIGCAP * 128;
constexpr int V3_SM_STAGE = V3_SM_LISTS + V3_LIST_ENTRIES * 2;
constexpr int V3_SM_KEYS = V3_SM_STAGE + V3_NW * V3_STAGE_WORDS * 4;
constexpr int V3_SM_SPAN_BYTES = V3_SM_KEYS - V3_SM_LISTS;  // 14336: 16 triangles x up to 448 columns x 2 bytes
// Visible-triangle flags of the fused G-buffer scan (non-depth shaders, see V3Vis): one bit per triangle, behind the
// span table (<= 16 x 255 x 2 bytes) in the lists + stage area, which is dead during the resolve.
constexpr int V3_SM_VISFLAGS = V3_SM_LISTS + 8192;
constexpr int V3_VIS_MAXT = (V3_SM_KEYS - V3_SM_VISFLAGS) * 8;   // 49152 triangles
static_assert(V3_BIGCAP * 255 * 2 <= 8192, "span table overlaps the visibility flags");
static_assert(V3_SM_KEYS % 16 == 0, "key tile must be 16-byte aligned");
static_assert(V3_BIGCAP * 255 * 2 <= V3_SM_SPAN_BYTES, "span table does not fit");
struct V3Layout { int keys, xs, ys, total; };
__host__ __device__ inline V3Layout v3_layout(int W, int H, int key_bytes) {
  V3Layout L;
  L.keys = V3_SM_KEYS;
  L.xs = L.keys + ((W * H * key_bytes + 15) & ~15);
  L.ys = L.xs + ((W * 4 + 15) & ~15);
  L.total = L.ys + ((H * 4 + 15) & ~15);
  return L;
}

// Large-triangle record in shared memory, every coefficient stored TWICE (a "splat" pair): one LDS.64 yields
// the 64-bit register pair the packed FMUL2 / FADD2 instructions of sm_100 take as an operand.
struct __align__(16) V3Big {  // 128 bytes
  float2 inv[9];
  float2 zc[3];
  int tri;
  short x0, x1, y0, y1;
  int pad[5];
};
static_assert(sizeof(V3Big) == 128, "V3Big must be 128 bytes");

// Fused scan of the G-buffer (what k_mark_visible does in a launch of its own, re-reading tri_id): the resolve of a
// non-depth pass flags every triangle it writes, and the CTA then emits the image's list of VISIBLE triangles
// (list[b][0..count[b])), their number and -- compact attribute records -- the triangle -> record slot map, for
// k_tri_attr / k_shade_rec.  All three NULL: no scan.
struct V3Vis { int* list; int* count; int* slot_map; };

// workspace of the single-tile kernel: one int32 spill slot per (image, triangle)
// (+ 512 slots so that the two CTAs of a clustered image -- V3_CLUSTER -- each own half of the image's row)
constexpr int V3_SPILL_PAD = 512;
__host__ inline size_t v3_workspace_bytes(int B, int T) { return (size_t)B * (size_t)((T > 0 ? T : 0) + V3_SPILL_PAD) * 4; }

// ---- thread-block cluster helpers (CLUSTER builds: one image split over the two CTAs of a cluster, small batches)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_map(uint32_t saddr, unsigned rank) {   // own shared address -> peer's
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ uint32_t ld_cluster_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_cluster_u32x4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared::cluster.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ unsigned atoms_add(uint32_t saddr, unsigned v) {
  unsigned old;  // one native shared-memory atomic (the CUDA intrinsic adds a warp-aggregation prologue)
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(saddr), "r"(v) : "memory");
  return old;
}

// per-image constants of the phase-A filter (shared memory)
struct V3Aux {
  float wa, wb;   // e_w   = wa * L1 + wb          (L1 = max over the triangle's vertices of |x|+|y|+|z|)
  float qa, qb;   // e_det = 3e-6 * (qa * L1 + qb)^3
  float mg_c;     // 3.8e-6 * view width  (bbox_margin's coefficient)
};

// packed fp32 add of sm_100 (FADD2): both halves rounded to nearest like two scalar FADDs
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }

__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Phase-A classification of one triangle from its three world-space vertices.  Returns -1 (certainly
// invisible) or the list (0..3) the triangle goes to.  Error bounds (u = 2^-24):
//   clip entry: |fma chain - reference order| <= 7.1 u S, S = sum |p_j m_j| + |m_3| <= rowmax * L1 + |m_3|
//   det:        12 u S' from the two evaluation orders + 3 * 7.1 u S'' from the perturbed entries, all
//               <= 3e-6 * (sum over rows of the S bound)^3
// The bbox uses the margin of bbox_margin (jr_device.cuh) + 1e-3 px, so that it contains the bbox the
// exact phase derives from the reference-order clip coordinates (they differ by < 1e-4 px on screen).
__device__ __forceinline__ int v3_filter(const float* __restrict__ m, const V3Aux& ax, float vp00, float vp03,
                                         float vp11, float vp13, const float p[9], int tw, int th) {
  float cx[3], cy[3], cw[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float x = p[3 * k], y = p[3 * k + 1], z = p[3 * k + 2];
    cx[k] = fmaf(x, m[0], fmaf(y, m[1], fmaf(z, m[2], m[3])));
    cy[k] = fmaf(x, m[4], fmaf(y, m[5], fmaf(z, m[6], m[7])));
    cw[k] = fmaf(x, m[12], fmaf(y, m[13], fmaf(z, m[14], m[15])));
  }
  const float l1 = fmaxf(fmaxf((fabsf(p[0]) + fabsf(p[1])) + fabsf(p[2]), (fabsf(p[3]) + fabsf(p[4])) + fabsf(p[5])),
                         (fabsf(p[6]) + fabsf(p[7])) + fabsf(p[8]));
  const float wthr = fmaf(l1, ax.wa, ax.wb);
  const float wmin = fminf(fminf(cw[0], cw[1]), cw[2]), wmax = fmaxf(fmaxf(cw[0], cw[1]), cw[2]);
  if (wmax < -wthr) return -1;  // certainly behind (all w <= 0)
  const float c0 = fmaf(cy[1], cw[2], -(cw[1] * cy[2]));
  const float c1 = fmaf(cw[1], cx[2], -(cx[1] * cw[2]));
  const float c2 = fmaf(cx[1], cy[2], -(cy[1] * cx[2]));
  const float det = fmaf(cx[0], c0, fmaf(cy[0], c1, cw[0] * c2));
  const float q = fmaf(l1, ax.qa, ax.qb);
  if (fmaf(3e-6f * q, q * q, det) < 1e-6f) return -1;  // certainly not (keep & front-facing); NaN stays
  int bw = tw, bh = th;
  if (wmin > wthr) {  // certainly in front: projected bbox
    const float r0 = rcp_fast(cw[0]), r1 = rcp_fast(cw[1]), r2 = rcp_fast(cw[2]);
    const float sx0 = fmaf(cx[0] * r0, vp00, vp03), sx1 = fmaf(cx[1] * r1, vp00, vp03), sx2 = fmaf(cx[2] * r2, vp00, vp03);
    const float sy0 = fmaf(cy[0] * r0, vp11, vp13), sy1 = fmaf(cy[1] * r1, vp11, vp13), sy2 = fmaf(cy[2] * r2, vp11, vp13);
    const float lox = fminf(fminf(sx0, sx1), sx2), hix = fmaxf(fmaxf(sx0, sx1), sx2);
    const float loy = fminf(fminf(sy0, sy1), sy2), hiy = fmaxf(fmaxf(sy0, sy1), sy2);
    float mg = 0.501f;
    if ((hix - lox) < 8.f && (hiy - loy) < 8.f) {
      const float ax_ = sx1 - sx0, ay_ = sy1 - sy0, bx_ = sx2 - sx0, by_ = sy2 - sy0, cx_ = sx2 - sx1, cy_ = sy2 - sy1;
      const float l2 = fmaxf(fmaxf(fmaf(ax_, ax_, ay_ * ay_), fmaf(bx_, bx_, by_ * by_)), fmaf(cx_, cx_, cy_ * cy_));
      const float area2 = fabsf(fmaf(ax_, by_, -(ay_ * bx_)));
      const float asp = l2 * rcp_fast(fmaxf(area2, 1e-20f));
      mg = fminf(0.5f, fmaxf(ax.mg_c * asp * asp, 0.015625f)) + 0.001f;
    }
    // fmaxf / fminf drop NaN towards "whole tile"
    const float mnx = fmaxf(lox - mg, 0.f), mxx = fminf(hix + mg, (float)(tw - 1));
    const float mny = fmaxf(loy - mg, 0.f), mxy = fminf(hiy + mg, (float)(th - 1));
    if (!(mnx <= mxx) || !(mny <= mxy)) return -1;
    const int x0 = (int)ceilf(mnx), x1 = (int)floorf(mxx), y0 = (int)ceilf(mny), y1 = (int)floorf(mxy);
    if (x0 > x1 || y0 > y1) return -1;
    bw = x1 - x0 + 1; bh = y1 - y0 + 1;
  }
  const int area = bw * bh;
  return area <= V3_A0 ? 0 : (area <= V3_A1 ? 1 : (area <= V3_A2 ? 2 : 3));
}

// stats[] slots (optional debug / measurement counters, JrRenderArgs.stats)
enum { V3S_TRIS = 0, V3S_PUSHED = 1, V3S_EXACT = 2, V3S_NTEST = 3, V3S_FRAGS = 4, V3S_ROUNDS = 5, V3S_BATCHES = 6, V3S_N = 8 };

// DepthShader quirk: a kept BACK-facing triangle 0 fills the pixels no candidate covers.  Exact
// arithmetic, once per image; kept out of line (the kernel's instruction footprint matters: the SM's warps sit in
// different phases and share the instruction cache).
__device__ __noinline__ void v3_tri0_setup(const float* __restrict__ w2c, const float* p, TriSetup* tri0, int* flag) {
  float M[9];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int rr = (r == 2) ? 3 : r;
    const float m0 = w2c[4 * rr], m1 = w2c[4 * rr + 1], m2 = w2c[4 * rr + 2], m3 = w2c[4 * rr + 3];
    M[0 + r] = ((p[0] * m0 + p[1] * m1) + p[2] * m2) + m3;
    M[3 + r] = ((p[3] * m0 + p[4] * m1) + p[5] * m2) + m3;
    M[6 + r] = ((p[6] * m0 + p[7] * m1) + p[8] * m2) + m3;
  }
  const float det = det3(M);
  const bool behind = (M[2] <= 0.f && M[5] <= 0.f && M[8] <= 0.f);
  if (det < -1e-6f && !behind) {
    float inv[9];
    lu_inverse3(M, inv);
    const float m0 = w2c[8], m1 = w2c[9], m2 = w2c[10], m3 = w2c[11];
#pragma unroll
    for (int k = 0; k < 9; ++k) tri0->inv[k] = inv[k];
    tri0->zc[0] = ((p[0] * m0 + p[1] * m1) + p[2] * m2) + m3;
    tri0->zc[1] = ((p[3] * m0 + p[4] * m1) + p[5] * m2) + m3;
    tri0->zc[2] = ((p[6] * m0 + p[7] * m1) + p[8] * m2) + m3;
    tri0->det = det;
    *flag = 1;
  }
}

// Scalar resolve: odd heights, unaligned buffers, or the DepthShader triangle-0 fallback.  Out of line (cold).
template <bool DEPTH, bool K32>
__device__ __noinline__ void v3_resolve_scalar(const unsigned long long* keys, const unsigned short* spans, const V3Big* bigq,
                                               const float* xs, const float* ys, int W, int H, int nbig, const TriSetup* tri0,
                                               float* __restrict__ z_out, int32_t* __restrict__ tri_out, float vp22,
                                               float vp23, int tid, float z_off, bool z_fill, float z_fillv,
                                               unsigned* vis_flags, uint32_t peer_keys_saddr = 0u, int part = 0,
                                               int nparts = 1) {
  typedef typename std::conditional<K32, uint32_t, unsigned long long>::type KeyT;
  const uint32_t* keys32 = reinterpret_cast<const uint32_t*>(keys);
  const int npix_img = W * H;
  const float rH = 1.0f / (float)H;
#pragma unroll 1
  for (int pix = tid + part * V3_THREADS; pix < npix_img; pix += V3_THREADS * nparts) {
    const int lx = (int)(((float)pix + 0.5f) * rH), ly = pix - lx * H;  // exact for pix < 2^16
    KeyT key = K32 ? (KeyT)keys32[pix] : (KeyT)keys[pix];
    if (K32 && peer_keys_saddr) key = (KeyT)min((uint32_t)key, ld_cluster_u32(peer_keys_saddr + 4u * pix));
#pragma unroll 1
    for (int e = 0; e < nbig; ++e) {
      const unsigned sp = spans[e * W + lx];
      if (ly < (int)(sp & 0xffu) || ly > (int)(sp >> 8)) continue;
      const V3Big& q = bigq[e];
      const float xn = xs[lx], yn = ys[ly];
      const float c0 = (xn * q.inv[0].x + yn * q.inv[3].x) + q.inv[6].x;
      const float c1 = (xn * q.inv[1].x + yn * q.inv[4].x) + q.inv[7].x;
      const float c2 = (xn * q.inv[2].x + yn * q.inv[5].x) + q.inv[8].x;
      const float z = (c0 * q.zc[0].x + c1 * q.zc[1].x) + c2 * q.zc[2].x;
      const float zw = z * vp22 + vp23;
      const KeyT kq = K32 ? (KeyT)min(orderable(zw), 0xFFFFFFFEu)
                          : (KeyT)(((unsigned long long)orderable(zw) << 32) | (unsigned)q.tri);
      if (kq < key) key = kq;
    }
    int tri = -1;
    const bool covered = key != (KeyT)~(KeyT)0;
    bool wrote = false;
    float zv = z_fillv;
    if (covered) {
      if (K32) {
        zv = from_orderable((uint32_t)key); wrote = true;
      } else {
        tri = (int)(unsigned)((unsigned long long)key & 0xFFFFFFFFull);
        if (DEPTH) { zv = from_orderable((uint32_t)((unsigned long long)key >> 32)); wrote = true; }
      }
    } else if (tri0) {
      float c[3];
      clip_coef(tri0->inv, xs[lx], ys[ly], c);
      if (c[0] >= 0.f && c[1] >= 0.f && c[2] >= 0.f) {
        const float z = (c[0] * tri0->zc[0] + c[1] * tri0->zc[1]) + c[2] * tri0->zc[2];
        zv = z * vp22 + vp23; wrote = true;
        tri = 0;
      }
    }
    if (DEPTH && (wrote || z_fill)) z_out[pix] = z_off != 0.f ? zv + z_off : zv;   // depth epilogue (jr_b200.h)
    if (tri_out) tri_out[pix] = tri;
    if (vis_flags && tri >= 0) {
      unsigned* w = vis_flags + (tri >> 5);
      const unsigned m = 1u << (tri & 31);
      if (!(*w & m)) atomicOr(w, m);
    }
  }
}

// out-of-line wrapper of the hierarchical warp raster (boxes of 256+ pixels: rare on the canvases this kernel serves)
template <bool K32>
__device__ __noinline__ void v3_raster_hier(float i0, float i1, float i2, float i3, float i4, float i5, float i6, float i7,
                                            float i8, float z0, float z1, float z2, unsigned tri, unsigned bb, int lane,
                                            const float* xs, const float* ys, uint32_t keys_saddr, int key_stride,
                                            float vp22, float vp23) {
  const float inv[9] = {i0, i1, i2, i3, i4, i5, i6, i7, i8};
  const float zc[3] = {z0, z1, z2};
  raster_hier_warp<K32>(inv, zc, tri, (int)(bb & 0xff), (int)((bb >> 16) & 0xff), (int)((bb >> 8) & 0xff), (int)(bb >> 24), lane,
                        xs, ys, keys_saddr, key_stride, vp22, vp23);
}

// INST: geometry instanced at the fetch (JrRenderArgs.inst_*); a template parameter so that the merged-array kernel
// does not carry the instancing code (the instruction footprint of this kernel is worth several percent).
// CLUSTER: the image is split over the TWO CTAs of a thread-block cluster (launch with cluster dimension 2, grid 2 B;
// z-only-key depth passes of batches too small to occupy the machine -- latency of one render, shadow passes of small
// batches).  Each CTA filters and rasterises every second group of 32 triangles into its OWN key tile; after a cluster
// barrier each resolves half of the pixel blocks from the minimum of both tiles (the peer's through distributed
// shared memory) and from the large triangles of both queues.  Same keys, same minima: bit-identical output.
template <bool DEPTH, bool K32, bool STATS, bool INST, bool CLUSTER = false>
__global__ void __launch_bounds__(V3_THREADS, K32 ? JR_V3_K32_CTAS : JR_V3_K64_CTAS)
k_vis3(const __grid_constant__ JrRenderArgs a, const V3Vis vis) {
  static_assert(DEPTH || !K32, "z-only keys are for the depth shader");
  static_assert(!CLUSTER || (K32 && !STATS), "the clustered build serves z-only-key depth passes");
  constexpr int NPART = CLUSTER ? 2 : 1;                 // CTAs per image
  constexpr int BIGCAP = CLUSTER ? V3_BIGCAP / 2 : V3_BIGCAP;   // per-CTA queue: the merged queue holds V3_BIGCAP
  extern __shared__ __align__(16) unsigned char smem[];
  const int W = a.W, H = a.H;
  const V3Layout L = v3_layout(W, H, K32 ? 4 : 8);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem + V3_SM_KEYS);
  uint32_t* keys32 = reinterpret_cast<uint32_t*>(smem + V3_SM_KEYS);
  float* xs = reinterpret_cast<float*>(smem + L.xs);
  float* ys = reinterpret_cast<float*>(smem + L.ys);
  V3Big* bigq = reinterpret_cast<V3Big*>(smem + V3_SM_BIGQ);
  unsigned short* lists = reinterpret_cast<unsigned short*>(smem + V3_SM_LISTS);
  unsigned short* spans = reinterpret_cast<unsigned short*>(smem + V3_SM_LISTS);  // resolve only
  __shared__ unsigned s_cnt01, s_cnt23;  // packed 16-bit push counters of lists (0, 1) and (2, 3)
  __shared__ unsigned s_head[5];  // consumption cursors of the four lists + the spill list
  __shared__ int bigq_n, tri0_flag, spill_n;
  __shared__ TriSetup tri0;
  __shared__ float s_w2c[16];
  __shared__ float s_vp[16];
  __shared__ V3Aux s_aux;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* stage = reinterpret_cast<float*>(smem + V3_SM_STAGE) + warp * V3_STAGE_WORDS;
  const int b = CLUSTER ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int crank = CLUSTER ? (int)(blockIdx.x & 1u) : 0;      // rank inside the cluster (cluster dimension (2, 1, 1))
  const uint32_t keys_saddr = (uint32_t)__cvta_generic_to_shared(keys);
  const uint32_t cnt01_saddr = (uint32_t)__cvta_generic_to_shared(&s_cnt01);
  const uint32_t cnt23_saddr = (uint32_t)__cvta_generic_to_shared(&s_cnt23);
  const uint32_t head_saddr = (uint32_t)__cvta_generic_to_shared(&s_head[0]);
  const int tw = W, th = H, tile_h = H;
  unsigned long long st_pushed = 0, st_exact = 0, st_ntest = 0, st_frags = 0, st_rounds = 0;

  if (tid < 16) {
    s_w2c[tid] = a.world_to_clip.ptr[(long long)b * a.world_to_clip.batch_stride + tid];
    s_vp[tid] = a.viewport.ptr[(long long)b * a.viewport.batch_stride + tid];
  }
  if (tid == 0) {
    bigq_n = 0; tri0_flag = 0; spill_n = 0; s_cnt01 = 0u; s_cnt23 = 0u;
    s_head[0] = s_head[1] = s_head[2] = s_head[3] = s_head[4] = 0u;
  }
  {
    const int n16 = (L.xs - V3_SM_KEYS) >> 4;
    uint4* k4 = reinterpret_cast<uint4*>(smem + V3_SM_KEYS);
#pragma unroll 2
    for (int i = tid; i < n16; i += V3_THREADS) k4[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
  }
  __syncthreads();
  // (no unrolling here and in the cold paths below: the kernel's instruction footprint is what the SM's
  // instruction cache holds for warps that sit in four different phases -- 6072 -> 4664 SASS instructions
  // was worth 12 % of the run time)
#pragma unroll 1
  for (int i = tid; i < tw; i += V3_THREADS) xs[i] = ((float)i - s_vp[3]) / s_vp[0];
#pragma unroll 1
  for (int i = tid; i < th; i += V3_THREADS) ys[i] = ((float)i - s_vp[7]) / s_vp[5];
  if (tid == 0) {
    const float* m = s_w2c;
    const float rx = fmaxf(fmaxf(fabsf(m[0]), fabsf(m[1])), fabsf(m[2]));
    const float ry = fmaxf(fmaxf(fabsf(m[4]), fabsf(m[5])), fabsf(m[6]));
    const float rw = fmaxf(fmaxf(fabsf(m[12]), fabsf(m[13])), fabsf(m[14]));
    s_aux.wa = 1e-6f * rw; s_aux.wb = 1e-6f * fabsf(m[15]);
    s_aux.qa = (rx + ry) + rw; s_aux.qb = (fabsf(m[3]) + fabsf(m[7])) + fabsf(m[15]);
    s_aux.mg_c = 3.8e-6f * (2.f * fmaxf(s_vp[0], s_vp[5]));
  }
  __syncthreads();

  const float vp00 = s_vp[0], vp03 = s_vp[3], vp11 = s_vp[5], vp13 = s_vp[7];
  const float vp22 = s_vp[10], vp23 = s_vp[11];
  const float* __restrict__ pos = a.position.ptr + (long long)b * a.position.batch_stride;
  const int32_t* __restrict__ faces = a.faces.ptr + (long long)b * a.faces.batch_stride;
  const int vmax = a.n_pos - 1;
  // (B, T + V3_SPILL_PAD) int32, see v3_workspace_bytes; a clustered image's two CTAs take half a row each
  int* __restrict__ spill = reinterpret_cast<int*>(a.workspace) + (long long)b * (a.T + V3_SPILL_PAD) +
                            (CLUSTER ? crank * ((a.T + V3_SPILL_PAD) >> 1) : 0);
  const unsigned lt_mask = (1u << lane) - 1u;

  // ------------------------------------------------------------------ exact rasterisers (k_vis2's)
  auto raster_small = [&](const float* inv, const float* zc, int tri, int x0, int x1, int y0, int y1) {
    for (int x = x0; x <= x1; ++x) {
      const float xn = xs[x];
      const float pk0 = xn * inv[0], pk1 = xn * inv[1], pk2 = xn * inv[2];
      for (int y = y0; y <= y1; ++y) {
        const float yn = ys[y];
        const float c0 = (pk0 + yn * inv[3]) + inv[6];
        const float c1 = (pk1 + yn * inv[4]) + inv[7];
        const float c2 = (pk2 + yn * inv[5]) + inv[8];
        if (STATS) ++st_ntest;
        if (c0 >= 0.f && c1 >= 0.f && c2 >= 0.f) {
          const float z = (c0 * zc[0] + c1 * zc[1]) + c2 * zc[2];
          const float zw = z * vp22 + vp23;
          if (STATS) ++st_frags;
          put_key<K32>(keys_saddr, x * tile_h + y, zw, (unsigned)tri);
        }
      }
    }
  };
  auto raster_flat = [&](const float* inv, const float* zc, int tri, int x0, int y0, int bw, int bh, int li, int nlanes) {
    const int n = bw * bh;
    const float rbh = 1.0f / (float)bh;
    for (int i = li; i < n; i += nlanes) {
      const int dx = (int)(((float)i + 0.5f) * rbh);
      const int x = x0 + dx, y = y0 + (i - dx * bh);
      const float xn = xs[x], yn = ys[y];
      const float c0 = (xn * inv[0] + yn * inv[3]) + inv[6];
      const float c1 = (xn * inv[1] + yn * inv[4]) + inv[7];
      const float c2 = (xn * inv[2] + yn * inv[5]) + inv[8];
      if (STATS) ++st_ntest;
      if (c0 >= 0.f && c1 >= 0.f && c2 >= 0.f) {
        const float z = (c0 * zc[0] + c1 * zc[1]) + c2 * zc[2];
        const float zw = z * vp22 + vp23;
        if (STATS) ++st_frags;
        put_key<K32>(keys_saddr, x * tile_h + y, zw, (unsigned)tri);
      }
    }
  };
  // exact set-up of the surviving lanes + dispatch by bbox size (warp-uniform call)
  auto fire = [&](bool valid, const float* M, const float* zc, int tri, unsigned bb, int small_limit) {
    float inv[9];
    const int x0 = bb & 0xff, x1 = (bb >> 8) & 0xff, y0 = (bb >> 16) & 0xff, y1 = bb >> 24;
    const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
    const int area = valid ? bw * bh : 0;
    if (valid) lu_inverse3(M, inv);
    bool is_medium = area > small_limit && area <= V2_MEDIUM_AREA;
    if (area > V2_MEDIUM_AREA) {
      const int slot = atomicAdd(&bigq_n, 1);
      if (slot < BIGCAP) {
        V3Big& q = bigq[slot];
#pragma unroll
        for (int k = 0; k < 9; ++k) q.inv[k] = make_float2(inv[k], inv[k]);
#pragma unroll
        for (int k = 0; k < 3; ++k) q.zc[k] = make_float2(zc[k], zc[k]);
        q.tri = tri;
        q.x0 = (short)x0; q.x1 = (short)x1; q.y0 = (short)y0; q.y1 = (short)y1;
      } else {
        is_medium = true;  // queue full: the warp takes it
      }
    }
    if (area > 0 && area <= small_limit) raster_small(inv, zc, tri, x0, x1, y0, y1);
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
      const int n = (sx1 - sx0 + 1) * (sy1 - sy0 + 1);
      if (n >= V2_HIER_AREA) {
        if (STATS && lane == 0) st_ntest += (unsigned long long)n;  // upper bound: the hierarchical sweep skips blocks
        v3_raster_hier<K32>(binv[0], binv[1], binv[2], binv[3], binv[4], binv[5], binv[6], binv[7], binv[8], bzc[0], bzc[1],
                            bzc[2], (unsigned)btri, bbb, lane, xs, ys, keys_saddr, tile_h, vp22, vp23);
      } else {
        raster_flat(binv, bzc, btri, sx0, sy0, sx1 - sx0 + 1, sy1 - sy0 + 1, lane, 32);
      }
    }
  };
  // reference-order (exact) clip transform, cull and bbox of triangle t: k_vis2's per-triangle body
  auto exact_setup = [&](int t, float* M, float* zc, unsigned& bb) -> bool {
    const int i0 = min(max(faces[3 * t + 0], 0), vmax), i1 = min(max(faces[3 * t + 1], 0), vmax),
              i2 = min(max(faces[3 * t + 2], 0), vmax);
    Vec3 q0{pos[3 * i0], pos[3 * i0 + 1], pos[3 * i0 + 2]}, q1{pos[3 * i1], pos[3 * i1 + 1], pos[3 * i1 + 2]},
        q2{pos[3 * i2], pos[3 * i2 + 1], pos[3 * i2 + 2]};
    if (INST && instanced(a)) {
      float pl[9] = {q0.x, q0.y, q0.z, q1.x, q1.y, q1.z, q2.x, q2.y, q2.z};
      instance_triangle(a, b, i0, i1, i2, pl);
      q0 = Vec3{pl[0], pl[1], pl[2]}; q1 = Vec3{pl[3], pl[4], pl[5]}; q2 = Vec3{pl[6], pl[7], pl[8]};
    }
    int x0, x1, y0, y1;
    if (!exact_cull_bbox(s_w2c, vp00, vp03, vp11, vp13, tw, th, q0, q1, q2, M, x0, x1, y0, y1)) return false;
    const float p0x = q0.x, p0y = q0.y, p0z = q0.z;
    const float p1x = q1.x, p1y = q1.y, p1z = q1.z;
    const float p2x = q2.x, p2y = q2.y, p2z = q2.z;
    const float m0 = s_w2c[8], m1 = s_w2c[9], m2 = s_w2c[10], m3 = s_w2c[11];
    zc[0] = ((p0x * m0 + p0y * m1) + p0z * m2) + m3;
    zc[1] = ((p1x * m0 + p1y * m1) + p1z * m2) + m3;
    zc[2] = ((p2x * m0 + p2y * m1) + p2z * m2) + m3;
    bb = (unsigned)x0 | ((unsigned)x1 << 8) | ((unsigned)y0 << 16) | ((unsigned)y1 << 24);
    return true;
  };

  // ================================================================== phase A: filter, lane = triangle
  // (static interleaved assignment of 32-triangle groups to warps: the filter costs the same for every group).
  // The face indices of a warp's NEXT group are loaded one iteration ahead, so a group waits for one global
  // round trip instead of two dependent ones.  (An L1 prefetch of the next group's vertex lines changed nothing:
  // with ~45 KB of shared memory per CTA the L1 keeps too few lines.)
  constexpr int GSTRIDE = V3_THREADS * NPART;   // triangles between a warp's consecutive groups
  const int g_first = (warp + V3_NW * crank) * 32;
  int nf0 = 0, nf1 = 0, nf2 = 0;
  if (g_first + lane < a.T) {
    const int tq = g_first + lane;
    nf0 = faces[3 * tq]; nf1 = faces[3 * tq + 1]; nf2 = faces[3 * tq + 2];
  }
  for (int t0 = g_first; t0 < a.T; t0 += GSTRIDE) {
    const int t = t0 + lane;
    const bool in = t < a.T;
    // out-of-range indices are clamped (the reference's gathers clamp too) -- never read out of bounds
    const int i0 = min(max(nf0, 0), vmax), i1 = min(max(nf1, 0), vmax), i2 = min(max(nf2, 0), vmax);
    if (t + GSTRIDE < a.T) {
      const int tq = t + GSTRIDE;
      nf0 = faces[3 * tq]; nf1 = faces[3 * tq + 1]; nf2 = faces[3 * tq + 2];
    }
    float p[9];
    // 32 triangles reading 96 consecutive vertices: nine coalesced loads through the warp's stage
    // instead of nine 36-byte-stride gathers (9 cache lines each)
    const int vbase = __shfl_sync(0xffffffffu, i0, 0);
    const bool seq = !in || (i0 == vbase + 3 * lane && i1 == i0 + 1 && i2 == i0 + 2);
    if (__all_sync(0xffffffffu, seq) && vbase + 96 <= a.n_pos) {
      const float* __restrict__ src = pos + 3 * vbase + lane;
#pragma unroll
      for (int k = 0; k < 9; ++k) stage[32 * k + lane] = src[32 * k];
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 9; ++k) p[k] = stage[9 * lane + k];
      __syncwarp();
    } else {
      p[0] = pos[3 * i0]; p[1] = pos[3 * i0 + 1]; p[2] = pos[3 * i0 + 2];
      p[3] = pos[3 * i1]; p[4] = pos[3 * i1 + 1]; p[5] = pos[3 * i1 + 2];
      p[6] = pos[3 * i2]; p[7] = pos[3 * i2 + 1]; p[8] = pos[3 * i2 + 2];
    }
    if (INST && instanced(a) && in) {
      // instanced geometry: local -> world (the merge's own arithmetic; the filter needs no more than that)
      instance_triangle(a, b, i0, i1, i2, p);
    }
    int cls = -1;
    if (in) cls = v3_filter(s_w2c, s_aux, vp00, vp03, vp11, vp13, p, tw, th);
    if (DEPTH && t == 0) {
      // (a copy: passing `p` itself to the out-of-line call would pin the array to local memory for every group)
      float p0[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) p0[k] = p[k];
      v3_tri0_setup(s_w2c, p0, &tri0, &tri0_flag);
    }
    // ---- push survivors: two native shared atomics reserve slots in the four lists
    const unsigned ms = __ballot_sync(0xffffffffu, cls >= 0);
    if (ms) {
      const unsigned b0 = __ballot_sync(0xffffffffu, cls >= 0 && (cls & 1));
      const unsigned b1 = __ballot_sync(0xffffffffu, cls >= 0 && (cls & 2));
      const unsigned m0 = ms & ~b0 & ~b1, m1 = ms & b0 & ~b1, m2 = ms & ~b0 & b1, m3 = ms & b0 & b1;
      unsigned base01 = 0, base23 = 0;
      if (lane == 0) {
        if (m0 | m1) base01 = atoms_add(cnt01_saddr, (unsigned)__popc(m0) | ((unsigned)__popc(m1) << 16));
        if (m2 | m3) base23 = atoms_add(cnt23_saddr, (unsigned)__popc(m2) | ((unsigned)__popc(m3) << 16));
      }
      base01 = __shfl_sync(0xffffffffu, base01, 0);
      base23 = __shfl_sync(0xffffffffu, base23, 0);
      if (cls >= 0) {
        // (arithmetic instead of per-lane selects: the lists are 1024 + 3 x 512 entries)
        static_assert(V3_CAP0 == 1024 && V3_CAP1 == 512 && V3_CAP2 == 512 && V3_CAP3 == 512, "list offsets below");
        const unsigned mine = ms & ((cls & 1) ? b0 : ~b0) & ((cls & 2) ? b1 : ~b1);
        const int off = (cls + (cls != 0)) << 9;                  // 0, 1024, 1536, 2048
        const int cap = 512 << (cls == 0);
        const unsigned bs = (cls & 2) ? base23 : base01;
        const int slot = (int)((bs >> ((cls & 1) << 4)) & 0xffffu) + __popc(mine & lt_mask);
        // list full (or a triangle index beyond 16 bits): the triangle goes to the image's spill list in the
        // workspace (global memory) and is set up after the lists
        if (slot < cap && t < V3_TMAX16) lists[off + slot] = (unsigned short)t;
        else spill[atomicAdd(&spill_n, 1)] = t;
        if (STATS) ++st_pushed;
      }
    }
  }
  __syncthreads();
  // ================================================================== phase C: exact, full warps
  {
    const unsigned c01 = s_cnt01, c23 = s_cnt23;
    const int n_sp = spill_n;
    // l = 4: the spill list (normally empty), then the lists from large to small boxes
#pragma unroll 1
    for (int l = n_sp > 0 ? 4 : 3; l >= 0; --l) {
      const int cap = l == 0 ? V3_CAP0 : V3_CAP1;
      const int n_l = l == 4 ? n_sp : min((int)(((l < 2 ? c01 : c23) >> (16 * (l & 1))) & 0xffffu), cap);
      const int off = l == 0 ? V3_OFF0 : (l == 1 ? V3_OFF1 : (l == 2 ? V3_OFF2 : V3_OFF3));
      const int small_limit = l == 2 ? V3_A2 : V2_SMALL_AREA;
      // entries per claim: the heavy lists are handed out in small portions so that all eight warps share
      // them (one warp sweeping every medium triangle of the image was the straggler at the barrier below)
      const int grain = l == 3 ? JR_V3_GRAIN3 : (l == 2 ? JR_V3_GRAIN2 : 32);
      for (;;) {
        unsigned ru = 0;
        if (lane == 0) ru = atoms_add(head_saddr + 4u * l, (unsigned)grain);
        const int r = (int)__shfl_sync(0xffffffffu, ru, 0);
        if (r >= n_l) break;
        if (STATS && lane == 0) ++st_rounds;
        const int e = r + lane;
        bool surv = false;
        float M[9], zc[3];
        unsigned bb = 0;
        int t = 0;
        if (lane < grain && e < n_l) {
          t = l == 4 ? spill[e] : (int)lists[off + e];
          surv = exact_setup(t, M, zc, bb);
          if (STATS && surv) ++st_exact;
        }
        if (__ballot_sync(0xffffffffu, surv)) fire(surv, M, zc, t, bb, small_limit);
      }
    }
  }
  uint32_t peer_keys_saddr = 0u;
  int nbig_merged = 0;
  if (CLUSTER) {
    // both key tiles and both queues are complete: merge the peer's large triangles (and the triangle-0 record, which
    // lives in rank 0) into this CTA's shared memory; the peer's KEYS are read in place by the resolve
    cluster_sync_all();
    const unsigned peer = (unsigned)(crank ^ 1);
    peer_keys_saddr = cluster_map(keys_saddr, peer);
    const int n_own = min(bigq_n, BIGCAP);
    const int n_peer = min((int)ld_cluster_u32(cluster_map((uint32_t)__cvta_generic_to_shared(&bigq_n), peer)), BIGCAP);
    const uint32_t peer_bigq = cluster_map((uint32_t)__cvta_generic_to_shared(bigq), peer);
    uint32_t* own_words = reinterpret_cast<uint32_t*>(bigq + n_own);
    for (int i = tid; i < n_peer * 32; i += V3_THREADS) own_words[i] = ld_cluster_u32(peer_bigq + 4u * i);
    if (crank == 1) {
      const uint32_t p_flag = cluster_map((uint32_t)__cvta_generic_to_shared(&tri0_flag), 0u);
      const uint32_t p_tri0 = cluster_map((uint32_t)__cvta_generic_to_shared(&tri0), 0u);
      uint32_t* t0w = reinterpret_cast<uint32_t*>(&tri0);
      for (int i = tid; i < (int)(sizeof(TriSetup) / 4); i += V3_THREADS) t0w[i] = ld_cluster_u32(p_tri0 + 4u * i);
      if (tid == 0) tri0_flag = (int)ld_cluster_u32(p_flag);
    }
    nbig_merged = n_own + n_peer;   // (bigq_n itself stays as it is: the peer reads it)
    __syncthreads();
  } else {
    __syncthreads();
  }

  // ================================================================== resolve
  const bool use0 = DEPTH && tri0_flag;
  const int npix_img = W * H;
  const int nbig = CLUSTER ? nbig_merged : min(bigq_n, V3_BIGCAP);
  // fused G-buffer scan (V3Vis): flag bits of the triangles this image's resolve writes
  const bool mark = !DEPTH && vis.list != nullptr;
  unsigned* vis_flags = reinterpret_cast<unsigned*>(smem + V3_SM_VISFLAGS);
  if (mark) {
    for (int i = tid; i < (a.T + 31) >> 5; i += V3_THREADS) vis_flags[i] = 0u;
    if (!nbig) __syncthreads();   // (with large triangles the barrier behind the span table does it)
  }
  // ---- span table of the large triangles (the ground plane of a Brax scene).  For a fixed column x every
  // edge function  fl(fl(pk + fl(yn * i)) + c)  is monotone in yn (each rounded op is), and yn = ys[y] is
  // monotone in y: the rows where edge k is >= 0 form a prefix or a suffix of the column, the rows inside
  // the triangle an interval.  One thread per (triangle, column) finds it by bisection with the very
  // expressions of the rasterisers -- exactly the set the per-pixel test accepts -- and the resolve then
  // evaluates depth only where a triangle is inside (lists + stage are dead: the table lives there).
  if (nbig) {
    for (int task = tid; task < nbig * W; task += V3_THREADS) {
      const int e = task / W, x = task - e * W;
      const V3Big& q = bigq[e];
      int lo = q.y0, hi = q.y1;
      bool empty = x < q.x0 || x > q.x1;
      if (!empty) {
        const float xn = xs[x];
#pragma unroll 1
        for (int k = 0; k < 3 && !empty; ++k) {
          const float pk = xn * q.inv[k].x, ik = q.inv[3 + k].x, ck = q.inv[6 + k].x;
          const bool in_lo = ((pk + ys[lo] * ik) + ck) >= 0.f;
          const bool in_hi = ((pk + ys[hi] * ik) + ck) >= 0.f;
          if (in_lo && in_hi) continue;
          if (!in_lo && !in_hi) { empty = true; break; }
          // boundary between l (inside(l) == in_lo) and h = l + 1 (inside(h) == in_hi): start from the analytic
          // root of the edge function (approximate arithmetic is fine here), then walk to the exact boundary
          // with the exact test -- monotonicity makes the walk terminate at the unique switch point
          const float yroot = (-(pk + ck) * rcp_fast(ik) * vp11 + vp13);
          int l = min(max((int)floorf(fminf(fmaxf(yroot, (float)lo), (float)hi)), lo), hi - 1);
          while (l > lo && (((pk + ys[l] * ik) + ck) >= 0.f) != in_lo) --l;
          while (l + 1 < hi && (((pk + ys[l + 1] * ik) + ck) >= 0.f) == in_lo) ++l;
          if (in_lo) hi = l; else lo = l + 1;
        }
      }
      spans[task] = empty ? (unsigned short)0x00ffu : (unsigned short)(lo | (hi << 8));
    }
    if (STATS && a.stats && tid == 0) atomicAdd(a.stats + V3S_NTEST, (unsigned long long)nbig * npix_img);
    __syncthreads();
  }
  constexpr int PX = K32 ? 4 : 2;  // pixels per thread of the vectorised resolve
  int32_t* __restrict__ tri_out = a.tri_id ? a.tri_id + (long long)b * W * H : nullptr;
  float* __restrict__ z_out = DEPTH ? a.zbuffer + (long long)b * W * H : nullptr;
  // vector stores need 16-byte (z, z-only keys) / 8-byte (z, tri_id) aligned images
  const bool aligned = ((DEPTH ? (reinterpret_cast<uintptr_t>(z_out) & (K32 ? 15 : 7)) : 0) == 0) &&
                       ((tri_out ? (reinterpret_cast<uintptr_t>(tri_out) & 7) : 0) == 0);
  const bool fused = !use0 && (H % PX) == 0 && aligned;
  // depth epilogue (jr_b200.h): + offset on every written depth, optional fill of the uncovered pixels
  const float z_off = DEPTH ? a.depth_offset : 0.f;
  const bool z_fill = DEPTH && a.depth_fill != 0;
  const float z_fillv = a.depth_fill_value;
  auto zfin = [&](float v) { return z_off != 0.f ? v + z_off : v; };
  typedef typename std::conditional<K32, uint32_t, unsigned long long>::type KeyT;
  if (fused) {
    // A warp resolves blocks of 8 columns x (4 * PX) rows: lane = (column, group of PX consecutive rows), one
    // 128-bit LDS of the PX keys, 128-bit stores.  The compact footprint keeps a warp on one side of the
    // ground plane's diagonal most of the time (a strip of columns always straddles it), so the depth of a
    // large triangle is evaluated once per block, not once per triangle.  (Packed FMUL2 feeding FADD2 is not
    // used: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under -fmad=false, which changes
    // bits.  Scalar products feeding packed sums are safe -- build.sh checks the SASS holds no FFMA2.)
    // Each thread owns NV vectors = PXT consecutive rows of one column (per-thread overheads -- span fetch, triangle
    // coefficients, index arithmetic -- are paid once per PXT pixels).
    constexpr int NV = JR_V3_NV, PXT = PX * NV, RU = JR_V3_RESOLVE_UNROLL;
    const int cx = lane & 7, gy = lane >> 3;
    const int nbx = (W + 7) >> 3, nby = (H + 4 * PXT - 1) / (4 * PXT);
    const float rnby = 1.0f / (float)nby;
    for (int blk = warp + V3_NW * crank; blk < nbx * nby; blk += V3_NW * NPART) {
      const int bx = (int)(((float)blk + 0.5f) * rnby), by = blk - bx * nby;  // exact quotient (blk < 2^16)
      const int x = bx * 8 + cx, y = by * (4 * PXT) + gy * PXT;
      const bool live = x < W && y < H;
      const int nrow = live ? min(PXT, H - y) : 0;     // rows of this thread inside the canvas (a multiple of PX)
      const int i = live ? (x * H + y) / PX : 0;      // index of the first vector
      KeyT k[PXT];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int iv = (v * PX < nrow) ? i + v : i;
        if (K32) {
          uint4 kk = reinterpret_cast<const uint4*>(keys32)[iv];
          if (CLUSTER) {
            const uint4 pk = ld_cluster_u32x4(peer_keys_saddr + 16u * (unsigned)iv);
            kk.x = min(kk.x, pk.x); kk.y = min(kk.y, pk.y); kk.z = min(kk.z, pk.z); kk.w = min(kk.w, pk.w);
          }
          k[v * PX + 0] = kk.x; k[v * PX + 1] = kk.y; k[v * PX + 2 % PX] = kk.z; k[v * PX + 3 % PX] = kk.w;
        } else {
          const ulonglong2 kk = reinterpret_cast<const ulonglong2*>(keys)[iv];
          k[v * PX + 0] = (KeyT)kk.x; k[v * PX + 1] = (KeyT)kk.y;
        }
      }
      if (nbig) {
        const int xc = live ? x : 0, yc = live ? y : 0;
        const float xn = xs[xc];
        float yn[PXT];
#pragma unroll
        for (int p = 0; p < PXT; ++p) yn[p] = ys[min(yc + p, H - 1)];
        for (int e = 0; e < nbig; ++e) {
          const unsigned sp = spans[e * W + xc];
          const int lo = (int)(sp & 0xffu), hi = (int)(sp >> 8);
          const bool overlap = live && hi >= yc && lo <= yc + nrow - 1;
          if (!__any_sync(0xffffffffu, overlap)) continue;
          const V3Big& q = bigq[e];
          const float i3 = q.inv[3].x, i4 = q.inv[4].x, i5 = q.inv[5].x, i6 = q.inv[6].x, i7 = q.inv[7].x, i8 = q.inv[8].x;
          const float px0 = xn * q.inv[0].x, px1 = xn * q.inv[1].x, px2 = xn * q.inv[2].x;
          const float z0 = q.zc[0].x, z1 = q.zc[1].x, z2 = q.zc[2].x;
          const unsigned tri = (unsigned)q.tri;
          const int rlo = overlap ? lo - yc : PXT, rhi = overlap ? hi - yc : -1;   // span in this thread's rows
#pragma unroll RU
          for (int h = 0; h < PXT / 2; ++h) {
            // Two rows at a time: scalar FMUL products, packed FADD2 sums (sm_100; each half is rounded like the
            // scalar FADD: the rasterisers' expressions, same bits).  Branch-free: evaluate, then select by the span.
            const float ya = yn[2 * h], yb = yn[2 * h + 1];
            const float2 c0 = add2(add2(make_float2(px0, px0), make_float2(ya * i3, yb * i3)), make_float2(i6, i6));
            const float2 c1 = add2(add2(make_float2(px1, px1), make_float2(ya * i4, yb * i4)), make_float2(i7, i7));
            const float2 c2 = add2(add2(make_float2(px2, px2), make_float2(ya * i5, yb * i5)), make_float2(i8, i8));
            const float2 z = add2(add2(make_float2(c0.x * z0, c0.y * z0), make_float2(c1.x * z1, c1.y * z1)),
                                  make_float2(c2.x * z2, c2.y * z2));
            const float2 zw = add2(make_float2(z.x * vp22, z.y * vp22), make_float2(vp23, vp23));
            const float zwp[2] = {zw.x, zw.y};
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const int p = 2 * h + jj;
              const bool inside = (p >= rlo) && (p <= rhi);
              KeyT key = K32 ? (KeyT)min(orderable(zwp[jj]), 0xFFFFFFFEu)
                             : (KeyT)(((unsigned long long)orderable(zwp[jj]) << 32) | tri);
              key = inside ? key : (KeyT)~(KeyT)0;
              k[p] = key < k[p] ? key : k[p];
            }
          }
        }
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (v * PX >= nrow) continue;
        const int iv = i + v;
        if (K32) {
          const uint32_t k0 = (uint32_t)k[v * PX], k1 = (uint32_t)k[v * PX + 1], k2 = (uint32_t)k[v * PX + 2 % PX],
                         k3 = (uint32_t)k[v * PX + 3 % PX];
          if (z_fill || (k0 != ~0u && k1 != ~0u && k2 != ~0u && k3 != ~0u)) {
            reinterpret_cast<float4*>(z_out)[iv] =
                make_float4(zfin(k0 != ~0u ? from_orderable(k0) : z_fillv), zfin(k1 != ~0u ? from_orderable(k1) : z_fillv),
                            zfin(k2 != ~0u ? from_orderable(k2) : z_fillv), zfin(k3 != ~0u ? from_orderable(k3) : z_fillv));
          } else {
            if (k0 != ~0u) z_out[4 * iv] = zfin(from_orderable(k0));
            if (k1 != ~0u) z_out[4 * iv + 1] = zfin(from_orderable(k1));
            if (k2 != ~0u) z_out[4 * iv + 2] = zfin(from_orderable(k2));
            if (k3 != ~0u) z_out[4 * iv + 3] = zfin(from_orderable(k3));
          }
        } else {
          const unsigned long long q0 = k[v * PX], q1 = k[v * PX + 1];
          const bool e0 = q0 == ~0ull, e1 = q1 == ~0ull;
          if (DEPTH) {
            if (z_fill || (!e0 && !e1)) {
              reinterpret_cast<float2*>(z_out)[iv] =
                  make_float2(zfin(e0 ? z_fillv : from_orderable((uint32_t)(q0 >> 32))),
                              zfin(e1 ? z_fillv : from_orderable((uint32_t)(q1 >> 32))));
            } else {
              if (!e0) z_out[2 * iv] = zfin(from_orderable((uint32_t)(q0 >> 32)));
              if (!e1) z_out[2 * iv + 1] = zfin(from_orderable((uint32_t)(q1 >> 32)));
            }
          }
          if (tri_out)
            reinterpret_cast<int2*>(tri_out)[iv] = make_int2(e0 ? -1 : (int)(unsigned)q0, e1 ? -1 : (int)(unsigned)q1);
          if (!DEPTH && mark) {
            // neighbouring rows mostly share their triangle; a plain load first: most pixels stop there
            const unsigned t0 = (unsigned)q0, t1 = (unsigned)q1;
            if (!e0) {
              unsigned* w = vis_flags + (t0 >> 5);
              const unsigned m = 1u << (t0 & 31);
              if (!(*w & m)) atomicOr(w, m);
            }
            if (!e1 && (e0 || t1 != t0)) {
              unsigned* w = vis_flags + (t1 >> 5);
              const unsigned m = 1u << (t1 & 31);
              if (!(*w & m)) atomicOr(w, m);
            }
          }
        }
      }
    }
  } else {
    v3_resolve_scalar<DEPTH, K32>(keys, spans, bigq, xs, ys, W, H, nbig, use0 ? &tri0 : nullptr, z_out, tri_out, vp22, vp23, tid,
                                  z_off, z_fill, z_fillv, mark ? vis_flags : nullptr, peer_keys_saddr, crank, NPART);
  }
  if (CLUSTER) cluster_sync_all();   // the peer may still be reading this CTA's key tile
  if (!DEPTH && mark) {
    // flags -> the image's visible-triangle list in ASCENDING triangle order (deterministic slots; the attribute
    // kernel then walks the index buffer forwards), count, slot map: every thread owns a contiguous run of flag words,
    // a block-wide exclusive scan of the per-thread counts gives its first slot
    __syncthreads();
    int* __restrict__ list = vis.list + (long long)b * a.T;
    int* __restrict__ smap = vis.slot_map ? vis.slot_map + (long long)b * a.T : nullptr;
    const int n_words = (a.T + 31) >> 5;
    const int per = (n_words + V3_THREADS - 1) / V3_THREADS;
    const int w0 = min(tid * per, n_words), w1 = min(w0 + per, n_words);
    int cnt = 0;
    for (int i = w0; i < w1; ++i) cnt += __popc(vis_flags[i]);
    int incl = cnt;   // inclusive scan inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    __shared__ int s_scan[V3_NW];
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    int base = incl - cnt;
    for (int w = 0; w < warp; ++w) base += s_scan[w];
    int slot = base;
    for (int i = w0; i < w1; ++i) {
      unsigned bits = vis_flags[i];
      while (bits) {
        const int tri = 32 * i + __ffs(bits) - 1;
        bits &= bits - 1;
        list[slot] = tri;
        if (smap) smap[tri] = slot;
        ++slot;
      }
    }
    if (tid == V3_THREADS - 1) vis.count[b] = slot;   // the last thread's end == the total
  }

  if (STATS && a.stats) {
    unsigned long long v[5] = {st_pushed, st_exact, st_ntest, st_frags, st_rounds};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      unsigned long long x = v[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      v[k] = x;
    }
    if (lane == 0) {
      atomicAdd(a.stats + V3S_PUSHED, v[0]); atomicAdd(a.stats + V3S_EXACT, v[1]); atomicAdd(a.stats + V3S_NTEST, v[2]);
      atomicAdd(a.stats + V3S_FRAGS, v[3]); atomicAdd(a.stats + V3S_ROUNDS, v[4]);
      if (warp == 0) { atomicAdd(a.stats + V3S_BATCHES, 1ull); atomicAdd(a.stats + V3S_TRIS, (unsigned long long)a.T); }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Audit of the two conservative culls (tests only, jr_debug_audit_cull): one warp per triangle
// brute-forces EVERY pixel of the canvas with the exact edge functions -- the reference's own test, pipeline.py:190 --
// and counts what the fast paths would have lost:
//   counters[0]  triangles the phase-A filter of k_vis3 drops although the exact cull keeps them
//   counters[1]  pixels an exactly-kept triangle covers OUTSIDE the bbox the exact phase rasterises (bbox_margin)
//   counters[2]  pixels covered by triangles the exact bbox cull rejects altogether
//   counters[3]  triangles audited that the exact cull keeps,   counters[4]  inside pixels seen in total
// All of [0..2] must read 0.
__global__ void __launch_bounds__(256) k_audit_cull(const __grid_constant__ JrRenderArgs a, unsigned long long* __restrict__ counters) {
  __shared__ float s_w2c[16], s_vp[16];
  __shared__ V3Aux s_aux;
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 16) {
    s_w2c[threadIdx.x] = a.world_to_clip.ptr[(long long)b * a.world_to_clip.batch_stride + threadIdx.x];
    s_vp[threadIdx.x] = a.viewport.ptr[(long long)b * a.viewport.batch_stride + threadIdx.x];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float* m = s_w2c;
    const float rx = fmaxf(fmaxf(fabsf(m[0]), fabsf(m[1])), fabsf(m[2]));
    const float ry = fmaxf(fmaxf(fabsf(m[4]), fabsf(m[5])), fabsf(m[6]));
    const float rw = fmaxf(fmaxf(fabsf(m[12]), fabsf(m[13])), fabsf(m[14]));
    s_aux.wa = 1e-6f * rw; s_aux.wb = 1e-6f * fabsf(m[15]);
    s_aux.qa = (rx + ry) + rw; s_aux.qb = (fabsf(m[3]) + fabsf(m[7])) + fabsf(m[15]);
    s_aux.mg_c = 3.8e-6f * (2.f * fmaxf(s_vp[0], s_vp[5]));
  }
  __syncthreads();
  const int t = blockIdx.x * 8 + warp;
  if (t >= a.T) return;
  const float vp00 = s_vp[0], vp03 = s_vp[3], vp11 = s_vp[5], vp13 = s_vp[7];
  const float* __restrict__ pos = a.position.ptr + (long long)b * a.position.batch_stride;
  const int32_t* __restrict__ faces = a.faces.ptr + (long long)b * a.faces.batch_stride;
  const int vmax = a.n_pos - 1;
  const int i0 = min(max(faces[3 * t], 0), vmax), i1 = min(max(faces[3 * t + 1], 0), vmax), i2 = min(max(faces[3 * t + 2], 0), vmax);
  const Vec3 q0 = fetch_position(a, b, pos, i0), q1 = fetch_position(a, b, pos, i1), q2 = fetch_position(a, b, pos, i2);
  const float p[9] = {q0.x, q0.y, q0.z, q1.x, q1.y, q1.z, q2.x, q2.y, q2.z};
  // single-tile canvases only: the filter's bbox lives on the whole canvas there
  const bool single = a.W <= 255 && a.H <= 255;
  const int cls = single ? v3_filter(s_w2c, s_aux, vp00, vp03, vp11, vp13, p, a.W, a.H) : 0;
  float M[9];
  int x0 = 0, x1 = -1, y0 = 0, y1 = -1;
  const bool kept = exact_cull_bbox(s_w2c, vp00, vp03, vp11, vp13, a.W, a.H, q0, q1, q2, M, x0, x1, y0, y1);
  // the reference's own predicate for this triangle, without any bbox
  const bool cand = det3(M) > 1e-6f;   // (exact_cull_bbox filled M before any early return)
  unsigned long long in_total = 0, out_of_box = 0;
  if (cand) {
    float inv[9];
    lu_inverse3(M, inv);
    for (int pix = lane; pix < a.W * a.H; pix += 32) {
      const int x = pix / a.H, y = pix - x * a.H;
      const float xn = ((float)x - vp03) / vp00, yn = ((float)y - vp13) / vp11;
      const float c0 = (xn * inv[0] + yn * inv[3]) + inv[6];
      const float c1 = (xn * inv[1] + yn * inv[4]) + inv[7];
      const float c2 = (xn * inv[2] + yn * inv[5]) + inv[8];
      if (c0 >= 0.f && c1 >= 0.f && c2 >= 0.f) {
        ++in_total;
        if (!kept || x < x0 || x > x1 || y < y0 || y > y1) ++out_of_box;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    in_total += __shfl_xor_sync(0xffffffffu, in_total, o);
    out_of_box += __shfl_xor_sync(0xffffffffu, out_of_box, o);
  }
  if (lane == 0) {
    if (kept && cls < 0) atomicAdd(counters + 0, 1ull);
    if (kept && out_of_box) atomicAdd(counters + 1, out_of_box);
    if (!kept && out_of_box) atomicAdd(counters + 2, out_of_box);
    if (kept) atomicAdd(counters + 3, 1ull);
    atomicAdd(counters + 4, in_total);
  }
}

}  // namespace jr
