// jr_forward.cu -- forward kernels + C-ABI entry points (include/jr_b200.h).
//
//   k_visibility<DEPTH>  one CTA per (image, screen tile).  Streams the image's
//                        triangles: vertex transform + PerPrimitive setup in
//                        registers (exact op order), cull, rasterise the
//                        clamped bounding box against a shared-memory tile of
//                        packed 64-bit (orderable z | triangle id) keys with
//                        atomicMin == the reference's "min depth, first index"
//                        argmin (shader.py:207-217).  Small boxes: one lane per
//                        triangle; large boxes: queued and rasterised by the
//                        whole CTA with lanes over pixels.  Resolve writes the
//                        triangle-id G-buffer (and z for the depth shader).
//   k_shade<SHADER>      one thread per pixel: recompute the chosen triangle's
//                        setup, perspective-correct interpolate + fragment + mix
//                        fused, write z / canvas where kept.
//
// Compiled with -fmad=false (see jr_device.cuh).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/jr_b200.h"
#include "jr_device.cuh"
#include "jr_common.cuh"
#include "jr_shade.cuh"
#include "jr_visibility.cuh"
#include "jr_tiled.cuh"
#include <stdlib.h>

namespace jr {

// ------------------------------------------------------------------ visibility
constexpr int VIS_THREADS = 256;
constexpr int BIGQ_CAP = 128;
constexpr int SMALL_AREA = 48;   // bbox pixels handled by the owning lane
constexpr unsigned long long EMPTY_KEY = 0xFFFFFFFFFFFFFFFFull;

struct __align__(16) BigTri {
  float inv[9];
  float zc[3];
  int tri;
  short x0, x1, y0, y1;  // inclusive, tile-local
  int pad;
};
static_assert(sizeof(BigTri) == 64, "BigTri must be 64 bytes");

struct VisSmemLayout {
  size_t keys, xs, ys, bigq, total;
};
__host__ __device__ inline VisSmemLayout vis_layout(int tile_w, int tile_h) {
  VisSmemLayout L;
  L.keys = 0;
  L.xs = (size_t)tile_w * tile_h * 8;
  L.ys = L.xs + (size_t)tile_w * 4;
  size_t e = L.ys + (size_t)tile_h * 4;
  L.bigq = (e + 15) & ~(size_t)15;
  L.total = L.bigq + (size_t)BIGQ_CAP * sizeof(BigTri);
  return L;
}

__device__ __forceinline__ void raster_pixel(const float* inv, const float* zc, float pk0, float pk1,
                                             float pk2, float yn, float vp22, float vp23, int tri,
                                             unsigned long long* slot) {
  float c0 = (pk0 + yn * inv[3]) + inv[6];
  float c1 = (pk1 + yn * inv[4]) + inv[7];
  float c2 = (pk2 + yn * inv[5]) + inv[8];
  if (c0 >= 0.f && c1 >= 0.f && c2 >= 0.f) {
    float z = (c0 * zc[0] + c1 * zc[1]) + c2 * zc[2];
    float zw = z * vp22 + vp23;
    unsigned long long key = ((unsigned long long)orderable(zw) << 32) | (unsigned)tri;
    if (key < *slot) atomicMin(slot, key);
  }
}

template <bool DEPTH>
__global__ void __launch_bounds__(VIS_THREADS)
k_visibility(const __grid_constant__ JrRenderArgs a, int tile_w, int tile_h, int tiles_x, int tiles_y) {
  extern __shared__ __align__(16) unsigned char smem[];
  const VisSmemLayout L = vis_layout(tile_w, tile_h);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem + L.keys);
  float* xs = reinterpret_cast<float*>(smem + L.xs);
  float* ys = reinterpret_cast<float*>(smem + L.ys);
  BigTri* bigq = reinterpret_cast<BigTri*>(smem + L.bigq);
  __shared__ int bigq_n;
  __shared__ int tri0_flag;
  __shared__ TriSetup tri0;
  __shared__ float s_w2c[16];
  __shared__ float s_vp[16];

  const int tid = threadIdx.x;
  const int tiles = tiles_x * tiles_y;
  const int b = blockIdx.x / tiles;
  const int tile = blockIdx.x - b * tiles;
  const int tx0 = (tile / tiles_y) * tile_w;
  const int ty0 = (tile % tiles_y) * tile_h;
  const int tw = min(tile_w, a.W - tx0);
  const int th = min(tile_h, a.H - ty0);

  if (tid < 16) {
    s_w2c[tid] = a.world_to_clip.ptr[(long long)b * a.world_to_clip.batch_stride + tid];
    s_vp[tid] = a.viewport.ptr[(long long)b * a.viewport.batch_stride + tid];
  }
  if (tid == 0) { bigq_n = 0; tri0_flag = 0; }
  for (int i = tid; i < tile_w * tile_h; i += VIS_THREADS) keys[i] = EMPTY_KEY;
  __syncthreads();
  // pixel -> NDC (pipeline.py:177)
  for (int i = tid; i < tw; i += VIS_THREADS) xs[i] = ((float)(tx0 + i) - s_vp[3]) / s_vp[0];
  for (int i = tid; i < th; i += VIS_THREADS) ys[i] = ((float)(ty0 + i) - s_vp[7]) / s_vp[5];
  __syncthreads();

  const float vp00 = s_vp[0], vp03 = s_vp[3], vp11 = s_vp[5], vp13 = s_vp[7];
  const float vp22 = s_vp[10], vp23 = s_vp[11];
  const float* __restrict__ pos = a.position.ptr + (long long)b * a.position.batch_stride;
  const int32_t* __restrict__ faces = a.faces.ptr + (long long)b * a.faces.batch_stride;
  const float fx_lo = (float)tx0, fx_hi = (float)(tx0 + tw - 1);
  const float fy_lo = (float)ty0, fy_hi = (float)(ty0 + th - 1);

  for (int t = tid; t < a.T; t += VIS_THREADS) {
    const int i0 = faces[3 * t + 0], i1 = faces[3 * t + 1], i2 = faces[3 * t + 2];
    float c0[4], c1[4], c2[4];
    to_clip(s_w2c, pos[3 * i0], pos[3 * i0 + 1], pos[3 * i0 + 2], c0);
    to_clip(s_w2c, pos[3 * i1], pos[3 * i1 + 1], pos[3 * i1 + 2], c1);
    to_clip(s_w2c, pos[3 * i2], pos[3 * i2 + 1], pos[3 * i2 + 2], c2);
    float M[9];
    const float det = tri_matrix(c0, c1, c2, M);
    // candidate <=> keep & front <=> |det| > 1e-6 & det >= 0   (pipeline.py:98-100, :232)
    const bool cand = det > 1e-6f;
    // DepthShader quirk: a kept back-facing triangle 0 is written
    // where no candidate exists (argmin of all-inf is index 0).
    const bool fallback0 = DEPTH && (t == 0) && (det < -1e-6f);
    if (!cand && !fallback0) continue;
    const float w0 = c0[3], w1 = c1[3], w2 = c2[3];
    if (w0 <= 0.f && w1 <= 0.f && w2 <= 0.f) continue;  // never inside (Q4)

    int x0 = 0, x1 = tw - 1, y0 = 0, y1 = th - 1;  // tile-local, inclusive
    if (w0 > 0.f && w1 > 0.f && w2 > 0.f && !fallback0) {
      const float sx0 = (c0[0] / w0) * vp00 + vp03, sx1 = (c1[0] / w1) * vp00 + vp03,
                  sx2 = (c2[0] / w2) * vp00 + vp03;
      const float sy0 = (c0[1] / w0) * vp11 + vp13, sy1 = (c1[1] / w1) * vp11 + vp13,
                  sy2 = (c2[1] / w2) * vp11 + vp13;
      // conservative +-0.5 px margin; fmaxf/fminf drop NaN towards "full tile"
      float mnx = fmaxf(fminf(fminf(sx0, sx1), sx2) - 0.5f, fx_lo);
      float mxx = fminf(fmaxf(fmaxf(sx0, sx1), sx2) + 0.5f, fx_hi);
      float mny = fmaxf(fminf(fminf(sy0, sy1), sy2) - 0.5f, fy_lo);
      float mxy = fminf(fmaxf(fmaxf(sy0, sy1), sy2) + 0.5f, fy_hi);
      if (!(mnx <= mxx) || !(mny <= mxy)) continue;
      x0 = (int)ceilf(mnx) - tx0; x1 = (int)floorf(mxx) - tx0;
      y0 = (int)ceilf(mny) - ty0; y1 = (int)floorf(mxy) - ty0;
      if (x0 > x1 || y0 > y1) continue;
    }
    float inv[9];
    lu_inverse3(M, inv);
    const float zc[3] = {c0[2], c1[2], c2[2]};
    if (fallback0) {
      if (DEPTH) {
#pragma unroll
        for (int k = 0; k < 9; ++k) tri0.inv[k] = inv[k];
        tri0.zc[0] = zc[0]; tri0.zc[1] = zc[1]; tri0.zc[2] = zc[2];
        tri0.det = det;
        tri0_flag = 1;
      }
      continue;
    }
    const int area = (x1 - x0 + 1) * (y1 - y0 + 1);
    bool inline_raster = area <= SMALL_AREA;
    if (!inline_raster) {
      const int slot = atomicAdd(&bigq_n, 1);
      if (slot < BIGQ_CAP) {
        BigTri& q = bigq[slot];
#pragma unroll
        for (int k = 0; k < 9; ++k) q.inv[k] = inv[k];
        q.zc[0] = zc[0]; q.zc[1] = zc[1]; q.zc[2] = zc[2];
        q.tri = t;
        q.x0 = (short)x0; q.x1 = (short)x1; q.y0 = (short)y0; q.y1 = (short)y1;
      } else {
        inline_raster = true;  // queue full: slow but correct
      }
    }
    if (inline_raster) {
      for (int x = x0; x <= x1; ++x) {
        const float xn = xs[x];
        const float pk0 = xn * inv[0], pk1 = xn * inv[1], pk2 = xn * inv[2];
        for (int y = y0; y <= y1; ++y)
          raster_pixel(inv, zc, pk0, pk1, pk2, ys[y], vp22, vp23, t, &keys[x * tile_h + y]);
      }
    }
  }
  __syncthreads();
  // Large triangles: hierarchical, exact.  fl(fl(xn*i0 + yn*i1) + i2) is monotone in
  // xn and in yn (every rounded op is monotone), so over a block of pixels each
  // fp32 edge value attains its max / min at one of the 4 block corners: a block
  // whose corner max is < 0 for some edge contains no inside pixel (skipped), one
  // whose corner min is >= 0 for all edges is fully inside (edge tests skipped).
  // One lane classifies one 4x8 block; surviving blocks are rasterised with one
  // lane per pixel.
  {
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NW = VIS_THREADS / 32;
    const int nbig = min(bigq_n, BIGQ_CAP);
    int unit = 0;  // (entry, round of 32 blocks) work units, dealt round-robin to the warps
    for (int e = 0; e < nbig; ++e) {
      const BigTri& q = bigq[e];
      const int bw = q.x1 - q.x0 + 1, bh = q.y1 - q.y0 + 1;
      const int nby = (bh + 7) >> 3;
      const int nblk = ((bw + 3) >> 2) * nby;
      for (int base = 0; base < nblk; base += 32, ++unit) {
        if (unit % NW != warp) continue;
        const int blk = base + lane;
        bool live = false, full = false;
        int xa = 0, ya = 0;
        if (blk < nblk) {
          const int bx = blk / nby, by = blk - bx * nby;
          xa = q.x0 + 4 * bx; ya = q.y0 + 8 * by;
          const int xb = min(xa + 3, (int)q.x1), yb = min(ya + 7, (int)q.y1);
          const float xna = xs[xa], xnb = xs[xb], yna = ys[ya], ynb = ys[yb];
          live = true; full = true;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float pa = xna * q.inv[k], pb = xnb * q.inv[k];
            const float qa = yna * q.inv[3 + k], qb = ynb * q.inv[3 + k];
            const float v0 = (pa + qa) + q.inv[6 + k], v1 = (pa + qb) + q.inv[6 + k];
            const float v2 = (pb + qa) + q.inv[6 + k], v3 = (pb + qb) + q.inv[6 + k];
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
          if (x <= q.x1 && y <= q.y1) {
            const float xn = xs[x], yn = ys[y];
            const float c0 = (xn * q.inv[0] + yn * q.inv[3]) + q.inv[6];
            const float c1 = (xn * q.inv[1] + yn * q.inv[4]) + q.inv[7];
            const float c2 = (xn * q.inv[2] + yn * q.inv[5]) + q.inv[8];
            if (((m_full >> j) & 1u) || (c0 >= 0.f && c1 >= 0.f && c2 >= 0.f)) {
              const float z = (c0 * q.zc[0] + c1 * q.zc[1]) + c2 * q.zc[2];
              const float zw = z * vp22 + vp23;
              const unsigned long long key = ((unsigned long long)orderable(zw) << 32) | (unsigned)q.tri;
              unsigned long long* slot = &keys[x * tile_h + y];
              if (key < *slot) atomicMin(slot, key);
            }
          }
        }
      }
    }
  }
  __syncthreads();
  // resolve (flat index walked without integer division: VIS_THREADS = dq*th + dr)
  int32_t* __restrict__ tri_out = a.tri_id ? a.tri_id + (long long)b * a.W * a.H : nullptr;
  float* __restrict__ z_out = DEPTH ? a.zbuffer + (long long)b * a.W * a.H : nullptr;
  const bool use0 = DEPTH && tri0_flag;
  {
    const int dq = VIS_THREADS / th, dr = VIS_THREADS - dq * th;
    int lx = tid / th, ly = tid - lx * th;
    for (; lx < tw; lx += dq, ly += dr) {
      if (ly >= th) { ly -= th; ++lx; if (lx >= tw) break; }
      const unsigned long long key = keys[lx * tile_h + ly];
      const long long pix = (long long)(tx0 + lx) * a.H + (ty0 + ly);
      int tri = -1;
      if (key != EMPTY_KEY) {
        tri = (int)(unsigned)(key & 0xFFFFFFFFull);
        if (DEPTH) z_out[pix] = from_orderable((uint32_t)(key >> 32));
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

// --------------------------------------------------------------------- shading
template <int SHADER>
__global__ void __launch_bounds__(256) k_shade(const __grid_constant__ JrRenderArgs a) {
  const long long npix = (long long)a.W * a.H;
  const long long total = npix * a.B;
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < total;
       gi += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(gi / npix);
    const int pix = (int)(gi - (long long)b * npix);
    const int tri = a.tri_id[gi];
    if (tri < 0) continue;
    const int x = pix / a.H, y = pix - x * a.H;
    Frag f;
    shade_pixel<SHADER>(a, b, x, y, tri, f);
    if (f.keep) {
      a.zbuffer[gi] = f.zw;
      float* o = a.canvas + gi * 3;
      o[0] = f.col[0]; o[1] = f.col[1]; o[2] = f.col[2];
    } else {
      a.tri_id[gi] = -1;
    }
  }
}

// One thread per triangle: vertex stage of the shader -> attribute record (large canvases).
template <int SHADER>
__global__ void __launch_bounds__(128) k_tri_attr(const __grid_constant__ JrRenderArgs a, float* __restrict__ attrs) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t >= a.T) return;
  Frag f;
  frag_vertex<SHADER>(a, b, t, f);
  attr_store<SHADER>(f, attrs + ((size_t)b * a.T + t) * TA_FLOATS);
}

// Pixel stage from attribute records (same arithmetic as k_shade, the per-triangle part is shared).
template <int SHADER>
__global__ void __launch_bounds__(256) k_shade_rec(const __grid_constant__ JrRenderArgs a,
                                                   const float* __restrict__ attrs) {
  const long long npix = (long long)a.W * a.H;
  const long long total = npix * a.B;
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < total;
       gi += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(gi / npix);
    const int pix = (int)(gi - (long long)b * npix);
    const int tri = a.tri_id[gi];
    if (tri < 0) continue;
    const int x = pix / a.H, y = pix - x * a.H;
    Frag f;
    attr_load<SHADER>(a, b, attrs + ((size_t)b * a.T + tri) * TA_FLOATS, f);
    frag_pixel<SHADER>(a, b, x, y, f);
    if (f.keep) {
      a.zbuffer[gi] = f.zw;
      float* o = a.canvas + gi * 3;
      o[0] = f.col[0]; o[1] = f.col[1]; o[2] = f.col[2];
    } else {
      a.tri_id[gi] = -1;
    }
  }
}

__global__ void k_add_scalar(float* data, long long n, float v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    data[i] = data[i] + v;
}

// canvas (B,W,H,3) fp32 -> (B,H,W,3) uint8, flipped vertically (utils.py:79-98)
__global__ void k_to_uint8_display(const float* __restrict__ canvas, uint8_t* __restrict__ out, int B, int W, int H) {
  __shared__ float tile[32][33 * 3];
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  const float* src = canvas + (long long)b * W * H * 3;
  uint8_t* dst = out + (long long)b * W * H * 3;
  // load: rows of x, contiguous over (y,c)
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int x = x0 + r;
    for (int j = threadIdx.x; j < 96; j += blockDim.x) {
      const int y = y0 + j / 3;
      tile[r][j] = (x < W && y < H) ? src[((long long)x * H + y) * 3 + (j % 3)] : 0.f;
    }
  }
  __syncthreads();
  // store: rows of display-y, contiguous over (x,c)
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int y = y0 + r;
    if (y >= H) continue;
    const int row = H - 1 - y;
    for (int j = threadIdx.x; j < 96; j += blockDim.x) {
      const int x = x0 + j / 3;
      if (x >= W) continue;
      float v = tile[j / 3][r * 3 + (j % 3)];
      v = fminf(fmaxf(v, 0.f), 1.f) * 255.f;
      dst[((long long)row * W + x) * 3 + (j % 3)] = (uint8_t)v;
    }
  }
}

}  // namespace jr

// ===================================================================== C ABI
using namespace jr;

static int check_common(const JrRenderArgs* a) {
  if (!a) return JR_ERR_NULL;
  if (a->shader < 0 || a->shader >= JR_NUM_SHADERS) return JR_ERR_SHADER;
  if (a->B <= 0 || a->W <= 0 || a->H <= 0 || a->T < 0 || a->n_pos < 0) return JR_ERR_DIMS;
  if (a->W > 32767 || a->H > 32767) return JR_ERR_DIMS;
  if (!a->world_to_clip.ptr || !a->viewport.ptr || !a->zbuffer) return JR_ERR_NULL;
  if (!a->tri_id && a->shader != JR_DEPTH) return JR_ERR_NULL;
  if (a->T > 0 && (!a->position.ptr || !a->faces.ptr)) return JR_ERR_NULL;
  const int s = a->shader;
  if (s != JR_DEPTH) {
    if (!a->canvas || !a->normal.ptr || !a->light_colour.ptr) return JR_ERR_NULL;
  }
  if (s == JR_GOURAUD && (!a->colour.ptr || !a->light_direction.ptr)) return JR_ERR_NULL;
  if (s >= JR_GOURAUD_TEXTURE) {
    if (!a->uv.ptr || !a->texture.ptr) return JR_ERR_NULL;
    if (a->tex_w <= 0 || a->tex_h <= 0) return JR_ERR_DIMS;
  }
  if ((s >= JR_GOURAUD_TEXTURE && s <= JR_PHONG_DARBOUX) && !a->light_direction.ptr) return JR_ERR_NULL;
  if (s >= JR_PHONG && !a->world_to_eye_norm.ptr) return JR_ERR_NULL;
  if (s == JR_PHONG_DARBOUX && (!a->normal_map.ptr || !a->id_to_face.ptr || !a->faces_indices.ptr))
    return JR_ERR_NULL;
  if (s >= JR_PHONG_REFLECTION) {
    if (!a->light_dir_eye.ptr || !a->ambient.ptr || !a->diffuse.ptr || !a->specular.ptr ||
        !a->specular_map.ptr || !a->texture_shape.ptr || !a->texture_index.ptr)
      return JR_ERR_NULL;
    if (a->spec_w <= 0 || a->spec_h <= 0 || a->n_objects <= 0) return JR_ERR_DIMS;
  }
  if (s == JR_PHONG_REFLECTION_SHADOW) {
    if (!a->shadow_map.ptr || !a->shadow_strength.ptr || !a->shadow_world_to_clip.ptr ||
        !a->shadow_viewport.ptr)
      return JR_ERR_NULL;
    if (a->shadow_w <= 0 || a->shadow_h <= 0) return JR_ERR_DIMS;
  }
  return JR_OK;
}

// Tile choice: whole canvas in one CTA when the key tile fits comfortably in
// shared memory (<= 96 KB -> e.g. 84x84, 110x110), else 64x64 tiles.
static void choose_tiles(int W, int H, int* tw, int* th, int* nx, int* ny) {
  // (tile-local coordinates are packed into 8 bits by k_vis2)
  if ((size_t)W * H * 8 <= 96 * 1024 && W <= 255 && H <= 255) { *tw = W; *th = H; *nx = 1; *ny = 1; return; }
  *tw = 64; *th = 64;
  if (W < 64) *tw = W;
  if (H < 64) *th = H;
  *nx = (W + *tw - 1) / *tw;
  *ny = (H + *th - 1) / *th;
}

extern "C" {

int jr_abi_version(void) { return JR_ABI_VERSION; }

const char* jr_strerror(int s) {
  switch (s) {
    case JR_OK: return "ok";
    case JR_ERR_NULL: return "a required pointer is NULL";
    case JR_ERR_DIMS: return "bad dimensions";
    case JR_ERR_SHADER: return "unknown shader id";
    case JR_ERR_WORKSPACE: return "workspace too small";
    case JR_ERR_UNSUPPORTED: return "unsupported combination";
    case JR_ERR_CUDA: return "CUDA launch failed";
    default: return "unknown status";
  }
}

// Forward scratch: [binned visibility: triangle records + tile bitmasks][shading attribute records]
struct FwdLayout { size_t tiled, attr_off, total; bool use_attr; };
static FwdLayout fwd_layout(const JrRenderArgs* a) {
  FwdLayout F{};
  int tw, th, nx, ny;
  choose_tiles(a->W, a->H, &tw, &th, &nx, &ny);
  F.tiled = (nx * ny == 1) ? 0 : tiled_layout(a->B, a->W, a->H, a->T).total;
  // per-triangle attribute records pay off when a triangle is shared by several pixels
  F.use_attr = a->shader != JR_DEPTH && a->shader != JR_PHONG_DARBOUX && a->T > 0 &&
               (long long)a->W * a->H >= 2LL * a->T && getenv("JR_NO_ATTR") == nullptr;
  F.attr_off = (F.tiled + 255) & ~(size_t)255;
  F.total = F.use_attr ? F.attr_off + (size_t)a->B * a->T * TA_FLOATS * 4 : F.tiled;
  return F;
}

size_t jr_workspace_bytes(const JrRenderArgs* a) {
  if (!a || a->B <= 0 || a->W <= 0 || a->H <= 0) return 0;
  return fwd_layout(a).total;
}

long long jr_launch_count(void) { return jr::g_launches.load(); }

int jr_render_forward(const JrRenderArgs* a, jr_stream_t stream_) {
  int st = check_common(a);
  if (st != JR_OK) return st;
  cudaStream_t stream = (cudaStream_t)stream_;
  int tw, th, nx, ny;
  choose_tiles(a->W, a->H, &tw, &th, &nx, &ny);
  const long long ctas = (long long)a->B * nx * ny;
  if (ctas > 2147483647LL) return JR_ERR_DIMS;
  static const bool use_v1 = getenv("JR_VIS_V1") != nullptr;  // A/B switch: first kernel version
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(k_visibility<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_visibility<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_vis2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_vis2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done = true;
  }
  const bool depth = a->shader == JR_DEPTH;
  static const bool no_bins = getenv("JR_NO_BINS") != nullptr;  // A/B switch: every tile CTA scans all triangles
  if (nx * ny > 1 && !use_v1 && !no_bins) {
    // two-level path: per-triangle records + per-tile bitmasks, then one CTA per (image, tile)
    const TiledLayout TLy = tiled_layout(a->B, a->W, a->H, a->T);
    if (!a->workspace || a->workspace_bytes < TLy.total) return JR_ERR_WORKSPACE;
    if (a->B > 65535) return JR_ERR_DIMS;
    char* ws = (char*)a->workspace;
    TriRecord* recs = (TriRecord*)(ws + TLy.rec);
    unsigned* masks = (unsigned*)(ws + TLy.mask);
    cudaMemsetAsync(masks, 0, (size_t)a->B * TLy.tiles * TLy.words * 4, stream);
    static bool attr2 = false;
    if (!attr2) {
      cudaFuncSetAttribute(k_raster_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      cudaFuncSetAttribute(k_raster_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      attr2 = true;
    }
    if (a->T > 0) {
      dim3 g1((a->T + 255) / 256, a->B);
      if (depth) k_setup_bin<true><<<g1, 256, 0, stream>>>(*a, recs, masks, TLy);
      else k_setup_bin<false><<<g1, 256, 0, stream>>>(*a, recs, masks, TLy);
      jr::g_launches++;
    }
    const long long ctas2 = (long long)a->B * TLy.tiles;
    if (ctas2 > 2147483647LL) return JR_ERR_DIMS;
    const size_t sm = tl_smem().total;
    if (depth) k_raster_tile<true><<<(unsigned)ctas2, TL_THREADS, sm, stream>>>(*a, recs, masks, TLy);
    else k_raster_tile<false><<<(unsigned)ctas2, TL_THREADS, sm, stream>>>(*a, recs, masks, TLy);
  } else if (use_v1) {
    const VisSmemLayout L = vis_layout(tw, th);
    if (depth) k_visibility<true><<<(unsigned)ctas, VIS_THREADS, L.total, stream>>>(*a, tw, th, nx, ny);
    else k_visibility<false><<<(unsigned)ctas, VIS_THREADS, L.total, stream>>>(*a, tw, th, nx, ny);
  } else {
    const V2Layout L = v2_layout(tw, th);
    if (depth) k_vis2<true><<<(unsigned)ctas, V2_THREADS, L.total, stream>>>(*a, tw, th, nx, ny);
    else k_vis2<false><<<(unsigned)ctas, V2_THREADS, L.total, stream>>>(*a, tw, th, nx, ny);
  }
  jr::g_launches++;
  if (!depth) {
    const long long total = (long long)a->B * a->W * a->H;
    const int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    if (blocks > 148LL * 64) blocks = 148LL * 64;
    const FwdLayout F = fwd_layout(a);
    if (F.use_attr) {
      if (!a->workspace || a->workspace_bytes < F.total) return JR_ERR_WORKSPACE;
      if (a->B > 65535) return JR_ERR_DIMS;
      float* attrs = (float*)((char*)a->workspace + F.attr_off);
      dim3 g1((a->T + 127) / 128, a->B);
#define JR_ATTR_CASE(S)                                                          \
  case S:                                                                        \
    k_tri_attr<S><<<g1, 128, 0, stream>>>(*a, attrs);                            \
    k_shade_rec<S><<<(unsigned)blocks, threads, 0, stream>>>(*a, attrs);         \
    break;
      switch (a->shader) {
        JR_ATTR_CASE(JR_GOURAUD)
        JR_ATTR_CASE(JR_GOURAUD_TEXTURE)
        JR_ATTR_CASE(JR_PHONG)
        JR_ATTR_CASE(JR_PHONG_REFLECTION)
        JR_ATTR_CASE(JR_PHONG_REFLECTION_SHADOW)
        default: return JR_ERR_SHADER;
      }
      jr::g_launches += 2;
    } else {
      switch (a->shader) {
        case JR_GOURAUD: k_shade<JR_GOURAUD><<<(unsigned)blocks, threads, 0, stream>>>(*a); break;
        case JR_GOURAUD_TEXTURE: k_shade<JR_GOURAUD_TEXTURE><<<(unsigned)blocks, threads, 0, stream>>>(*a); break;
        case JR_PHONG: k_shade<JR_PHONG><<<(unsigned)blocks, threads, 0, stream>>>(*a); break;
        case JR_PHONG_DARBOUX: k_shade<JR_PHONG_DARBOUX><<<(unsigned)blocks, threads, 0, stream>>>(*a); break;
        case JR_PHONG_REFLECTION: k_shade<JR_PHONG_REFLECTION><<<(unsigned)blocks, threads, 0, stream>>>(*a); break;
        case JR_PHONG_REFLECTION_SHADOW:
          k_shade<JR_PHONG_REFLECTION_SHADOW><<<(unsigned)blocks, threads, 0, stream>>>(*a); break;
        default: return JR_ERR_SHADER;
      }
      jr::g_launches++;
    }
  }
  return cudaGetLastError() == cudaSuccess ? JR_OK : JR_ERR_CUDA;
}

#define JR_SHADER_ENTRY(name, id)                                   \
  int name(const JrRenderArgs* a, jr_stream_t s) {                  \
    if (!a) return JR_ERR_NULL;                                     \
    if (a->shader != id) return JR_ERR_SHADER;                      \
    return jr_render_forward(a, s);                                 \
  }
JR_SHADER_ENTRY(jr_depth_forward, JR_DEPTH)
JR_SHADER_ENTRY(jr_gouraud_forward, JR_GOURAUD)
JR_SHADER_ENTRY(jr_gouraud_texture_forward, JR_GOURAUD_TEXTURE)
JR_SHADER_ENTRY(jr_phong_forward, JR_PHONG)
JR_SHADER_ENTRY(jr_phong_darboux_forward, JR_PHONG_DARBOUX)
JR_SHADER_ENTRY(jr_phong_reflection_forward, JR_PHONG_REFLECTION)
JR_SHADER_ENTRY(jr_phong_reflection_shadow_forward, JR_PHONG_REFLECTION_SHADOW)

int jr_add_scalar(float* data, long long n, float value, jr_stream_t stream) {
  if (!data) return JR_ERR_NULL;
  if (n <= 0) return JR_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_add_scalar<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(data, n, value);
  jr::g_launches++;
  return cudaGetLastError() == cudaSuccess ? JR_OK : JR_ERR_CUDA;
}

int jr_canvas_to_uint8_display(const float* canvas, uint8_t* out, int B, int W, int H, jr_stream_t stream) {
  if (!canvas || !out) return JR_ERR_NULL;
  if (B <= 0 || W <= 0 || H <= 0 || B > 65535) return JR_ERR_DIMS;
  dim3 grid((W + 31) / 32, (H + 31) / 32, B), block(32, 8);
  k_to_uint8_display<<<grid, block, 0, (cudaStream_t)stream>>>(canvas, out, B, W, H);
  jr::g_launches++;
  return cudaGetLastError() == cudaSuccess ? JR_OK : JR_ERR_CUDA;
}

}  // extern "C"
