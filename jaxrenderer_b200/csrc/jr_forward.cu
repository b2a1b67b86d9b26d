// jr_forward.cu -- forward launch logic + C-ABI entry points (include/jr_b200.h).
//
//   visibility  k_vis2 (jr_visibility.cuh): one CTA per image when the canvas fits one
//               shared-memory tile; k_setup_bin + k_raster_tile (jr_tiled.cuh) otherwise.
//               Both produce the triangle-id G-buffer (and z for the depth shader).
//   shading     k_shade<SHADER>: one thread per pixel, vertex + pixel stage recomputed per pixel
//               (Darboux shader);  k_mark_visible + k_tri_attr<SHADER> + k_shade_rec<SHADER>: vertex stage once
//               per VISIBLE triangle into an attribute record shared by the triangle's pixels (all other
//               shaders).  Same arithmetic either way (jr_shade.cuh).
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
#include "jr_vis3.cuh"
#include "jr_tiled.cuh"
#include <stdlib.h>
#include <mutex>

namespace jr {

// --------------------------------------------------------------------- shading
#ifndef JR_SHADE_CTAS
#define JR_SHADE_CTAS 4
#endif
template <int SHADER>
__global__ void __launch_bounds__(256, JR_SHADE_CTAS) k_shade(const __grid_constant__ JrRenderArgs a) {
  // grid = (blocks per image, images): all index arithmetic stays 32-bit (a 64-bit division per
  // pixel cost ~100 instructions)
  const int npix = a.W * a.H;
  for (int b = blockIdx.y; b < a.B; b += gridDim.y)
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
    const long long gi = (long long)b * npix + pix;
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
// Records are staged in shared memory and written out with coalesced 128-bit stores (a thread
// writing its own 176-byte record directly touches 44 different cache lines per warp store).
template <int SHADER>
__global__ void __launch_bounds__(128) k_tri_attr(const __grid_constant__ JrRenderArgs a, float* __restrict__ attrs,
                                                  const int* __restrict__ list, const int* __restrict__ count,
                                                  int rec_stride, bool compact, PixConst* __restrict__ pcs) {
  // records are staged in shared memory and written out as coalesced float4 (a thread writing its
  // own 176 B record with scalar stores made this kernel 4x slower)
  __shared__ __align__(16) float stage[128 * TA_FLOATS];
  __shared__ int s_tri[128];
  const int b = blockIdx.y;
  // the image's pixel-stage constants (read by k_shade_rec*), by the first warp of the image's first block
  if (blockIdx.x == 0 && threadIdx.x < 32)
    pix_const_write<SHADER>(a, b, pcs + b, reinterpret_cast<float*>(pcs + a.B) + (size_t)b * (a.W + a.H));
  // The number of visible triangles is only known on the device (a few hundred of thousands, typically): a small
  // grid of blocks per image walks the list in strides instead of ceil(T / 128) blocks of which most would find
  // nothing to do (106 k launched for 16 k with work on the facade workload).
  const int n_vis = count[b];
  constexpr int Q = TA_FLOATS / 4;
  const float4* src = reinterpret_cast<const float4*>(stage);
  float4* base = reinterpret_cast<float4*>(attrs + (size_t)b * rec_stride * TA_FLOATS);
  for (int i0 = blockIdx.x * 128; i0 < n_vis; i0 += gridDim.x * 128) {
    const int i = i0 + threadIdx.x;
    int t = -1;
    if (i < n_vis) {
      t = list[(long long)b * a.T + i];
      Frag f;
      frag_vertex<SHADER>(a, b, t, f);
      attr_store<SHADER>(f, stage + threadIdx.x * TA_FLOATS);
    }
    // record slot: the triangle id, or (compact layout, more triangles than pixels) the list position
    s_tri[threadIdx.x] = compact ? i : t;
    __syncthreads();
    const int n = min(128, n_vis - i0) * Q;
    for (int j = threadIdx.x; j < n; j += 128) {
      const int r = j / Q;
      base[(size_t)s_tri[r] * Q + (j - r * Q)] = src[j];
    }
    __syncthreads();   // the stage is reused by the next stride
  }
}

template <int SHADER>
__global__ void __launch_bounds__(256, JR_SHADE_CTAS) k_shade_rec(const __grid_constant__ JrRenderArgs a,
                                                   const float* __restrict__ attrs, int rec_stride,
                                                   const int* __restrict__ slot_map, const PixConst* __restrict__ pcs) {
  // One pixel per thread, no loops: grid = (ceil(W * H / 256), min(B, 65535), ceil(B / 65535)); all index arithmetic
  // stays 32-bit (a 64-bit division per pixel cost ~100 instructions); per-image constants through `pcs`
  // (PixConst, written by k_tri_attr); pix -> (x, y) by a multiply-high with the stored reciprocal of H (the integer
  // division was 3 % of the kernel).
  // (Colour stores stay three scalar stores per pixel: a warp's three STG cover 384 contiguous bytes, every sector
  // fully written.  Staging the warp's 96 floats in shared memory for 24 x 128-bit stores was measured slower --
  // 914 vs 868 us on the facade workload: five more instructions per pixel in an issue-bound kernel.)
  const int npix = a.W * a.H;
  const int b = blockIdx.z * 65535 + blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B || pix >= npix) return;
  const long long gi = (long long)b * npix + pix;
  const int tri = a.tri_id[gi];
  if (tri < 0) return;
  const PixConst* __restrict__ pc = pcs + b;
  const unsigned hm = pc->h_magic;
  const int x = hm ? (int)__umulhi((unsigned)pix, hm) : pix / a.H, y = pix - x * a.H;
  Frag f;
  const int slot = slot_map ? slot_map[(long long)b * a.T + tri] : tri;
  attr_load<SHADER, true>(a, b, attrs + ((size_t)b * rec_stride + slot) * TA_FLOATS, f);
  frag_pixel<SHADER, true>(a, b, x, y, f, pc);
  if (f.keep) {
    a.zbuffer[gi] = f.zw;
    float* o = a.canvas + gi * 3;
    o[0] = f.col[0]; o[1] = f.col[1]; o[2] = f.col[2];
  } else {
    a.tri_id[gi] = -1;
  }
}

// Shading with the display epilogue fused into the store (JrRenderArgs.canvas_u8): one CTA per
// 32 x 32 pixel tile of one image.  Threads read the G-buffer along y (coalesced), shade, write z, park the uint8
// colour of EVERY pixel of the tile in shared memory, and the tile is then written out along x -- the display
// layout (B, H, W, 3), flipped vertically -- with coalesced byte stores.  The fp32 canvas is never written.
template <int SHADER>
__global__ void __launch_bounds__(256, JR_SHADE_CTAS) k_shade_rec_u8(const __grid_constant__ JrRenderArgs a,
                                                      const float* __restrict__ attrs, int rec_stride,
                                                      const int* __restrict__ slot_map, int tiles_x, int tiles_y,
                                                      const PixConst* __restrict__ pcs) {
  __shared__ uint8_t tile[32][32 * 3 + 4];   // [y][x * 3 + c]
  const int b = blockIdx.y;
  const PixConst* __restrict__ pc = pcs + b;
  const int tx = blockIdx.x / tiles_y, ty = blockIdx.x - tx * tiles_y;
  const int x0 = tx * 32, y0 = ty * 32;
  const int ly = threadIdx.x & 31;
  const long long img = (long long)b * a.W * a.H;
#pragma unroll 1
  for (int lx = threadIdx.x >> 5; lx < 32; lx += 8) {
    const int x = x0 + lx, y = y0 + ly;
    if (x >= a.W || y >= a.H) continue;
    const long long gi = img + (long long)x * a.H + y;
    const int tri = a.tri_id[gi];
    float col[3] = {a.canvas_u8_background[0], a.canvas_u8_background[1], a.canvas_u8_background[2]};
    if (a.canvas) { col[0] = a.canvas[gi * 3]; col[1] = a.canvas[gi * 3 + 1]; col[2] = a.canvas[gi * 3 + 2]; }
    if (tri >= 0) {
      Frag f;
      const int slot = slot_map ? slot_map[(long long)b * a.T + tri] : tri;
      attr_load<SHADER, true>(a, b, attrs + ((size_t)b * rec_stride + slot) * TA_FLOATS, f);
      frag_pixel<SHADER, true>(a, b, x, y, f, pc);
      if (f.keep) {
        a.zbuffer[gi] = f.zw;
        col[0] = f.col[0]; col[1] = f.col[1]; col[2] = f.col[2];
      } else {
        a.tri_id[gi] = -1;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) tile[ly][lx * 3 + c] = (uint8_t)(fminf(fmaxf(col[c], 0.f), 1.f) * 255.f);
  }
  __syncthreads();
  const int tw = min(32, a.W - x0), th = min(32, a.H - y0);
  uint8_t* __restrict__ out = a.canvas_u8 + (long long)b * a.W * a.H * 3;
  for (int j = threadIdx.x; j < th * tw * 3; j += 256) {
    const int r = j / (tw * 3), q = j - r * (tw * 3);
    out[((long long)(a.H - 1 - (y0 + r)) * a.W + x0) * 3 + q] = tile[r][q];
  }
}

// ---------------------------------------------------------------- merge_objects (model.py:447-555)
__global__ void __launch_bounds__(256) k_merge_verts(const __grid_constant__ JrMergeArgs m) {
  // (staging the block's 3 KB in shared memory for 128-bit stores was measured slower: 292 vs 231 us)
  const int b = blockIdx.y;
  const int v = blockIdx.x * 256 + threadIdx.x;
  if (v >= m.n_verts) return;
  const float* lv = m.local_verts.ptr + (long long)b * m.local_verts.batch_stride + 3 * v;
  const int o = (m.vert_object.ptr + (long long)b * m.vert_object.batch_stride)[v];
  const float* s = m.scaling.ptr + (long long)b * m.scaling.batch_stride + 3 * o;
  const float* T = m.transform.ptr + (long long)b * m.transform.batch_stride + 16 * o;
  const float x = lv[0] * s[0], y = lv[1] * s[1], z = lv[2] * s[2];
  float h[4];
  to_clip(T, x, y, z, h);  // to_homogeneous(p) @ T^T
  float* out = m.out_verts + ((long long)b * m.n_verts + v) * 3;
  to_cartesian3(h, out[0], out[1], out[2]);  // geometry.py:183-202
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  // fixed-shape reduction: butterfly inside each warp, warps in order
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w];
  __syncthreads();
  return t;
}

// one CTA per (object, group of MN_GROUP batch elements): Camera.apply_vec with whole-array normalisation.
// The first Frobenius norm depends only on the object's local normals: computed once per CTA when the
// local mesh is shared by the batch (the usual case), and reused for the group's batch elements.
constexpr int MN_GROUP = 16;
// out_scales != nullptr: write only the two normalisation constants (f1, f2) of every (image, object) --
// jr_instance_norm_scales, for geometry instanced inside the render kernels -- and no normal.
__global__ void __launch_bounds__(256) k_merge_norms(const __grid_constant__ JrMergeArgs m, float* __restrict__ out_scales) {
  __shared__ float red[8];
  const int o = blockIdx.x;
  const bool shared_mesh = m.local_norms.batch_stride == 0 && m.norm_start.batch_stride == 0;
  float f1 = 0.f;
  if (out_scales && shared_mesh) {
    // Scales only, shared local mesh (instanced rendering: the per-step call of the Brax facade).  The once-normalised
    // normals n / f1 do not depend on the image: they are computed ONCE per CTA and kept in registers (3 IEEE
    // divisions per normal and image saved); only the rotation and the second norm run per image.
    const int n0 = m.norm_start.ptr[o], n1 = m.norm_start.ptr[o + 1];
    const float* __restrict__ ln = m.local_norms.ptr;
    float ss = 0.f;
    for (int i = n0 + threadIdx.x; i < n1; i += 256)
      ss += dot3(ln[3 * i], ln[3 * i + 1], ln[3 * i + 2], ln[3 * i], ln[3 * i + 1], ln[3 * i + 2]);
    f1 = sqrtf(block_sum_256(ss, red));
    constexpr int KEEP = 4;
    float nx[KEEP], ny[KEEP], nz[KEEP];
#pragma unroll
    for (int kk = 0; kk < KEEP; ++kk) {
      const int i = n0 + threadIdx.x + 256 * kk;
      nx[kk] = ny[kk] = nz[kk] = 0.f;
      if (i < n1) { nx[kk] = fdiv_z(ln[3 * i], f1); ny[kk] = fdiv_z(ln[3 * i + 1], f1); nz[kk] = fdiv_z(ln[3 * i + 2], f1); }
    }
    for (int b = blockIdx.y * MN_GROUP; b < min(m.B, (blockIdx.y + 1) * MN_GROUP); ++b) {
      const float* __restrict__ R = m.normal_matrix.ptr + (long long)b * m.normal_matrix.batch_stride + 16 * o;
      const float r0 = R[0], r1 = R[1], r2 = R[2], r4 = R[4], r5 = R[5], r6 = R[6], r8 = R[8], r9 = R[9], r10 = R[10];
      float ss2 = 0.f;
#pragma unroll
      for (int kk = 0; kk < KEEP; ++kk) {
        if (n0 + threadIdx.x + 256 * kk < n1) {   // same accumulation order as the general path below
          const float tx = (nx[kk] * r0 + ny[kk] * r1) + nz[kk] * r2, ty = (nx[kk] * r4 + ny[kk] * r5) + nz[kk] * r6,
                      tz = (nx[kk] * r8 + ny[kk] * r9) + nz[kk] * r10;
          ss2 += dot3(tx, ty, tz, tx, ty, tz);
        }
      }
      for (int i = n0 + threadIdx.x + 256 * KEEP; i < n1; i += 256) {
        const float x = fdiv_z(ln[3 * i], f1), y = fdiv_z(ln[3 * i + 1], f1), z = fdiv_z(ln[3 * i + 2], f1);
        const float tx = (x * r0 + y * r1) + z * r2, ty = (x * r4 + y * r5) + z * r6, tz = (x * r8 + y * r9) + z * r10;
        ss2 += dot3(tx, ty, tz, tx, ty, tz);
      }
      const float f2 = sqrtf(block_sum_256(ss2, red));
      if (threadIdx.x == 0) {
        out_scales[((long long)b * m.n_objects + o) * 2] = f1;
        out_scales[((long long)b * m.n_objects + o) * 2 + 1] = f2;
      }
    }
    return;
  }
  for (int b = blockIdx.y * MN_GROUP; b < min(m.B, (blockIdx.y + 1) * MN_GROUP); ++b) {
    const int32_t* ns = m.norm_start.ptr + (long long)b * m.norm_start.batch_stride;
    const int n0 = ns[o], n1 = ns[o + 1];
    const float* ln = m.local_norms.ptr + (long long)b * m.local_norms.batch_stride;
    const float* R = m.normal_matrix.ptr + (long long)b * m.normal_matrix.batch_stride + 16 * o;
    if (!shared_mesh || b == blockIdx.y * MN_GROUP) {
      float ss = 0.f;
      for (int i = n0 + threadIdx.x; i < n1; i += 256)
        ss += dot3(ln[3 * i], ln[3 * i + 1], ln[3 * i + 2], ln[3 * i], ln[3 * i + 1], ln[3 * i + 2]);
      f1 = sqrtf(block_sum_256(ss, red));
    }
    // rotated normals of this thread stay in registers between the second reduction and the final
    // division (objects of up to 256 * KEEP normals; larger ones recompute)
    constexpr int KEEP = 4;
    float keep[KEEP][3];
    float ss2 = 0.f;
#pragma unroll
    for (int kk = 0; kk < KEEP; ++kk) {
      const int i = n0 + threadIdx.x + 256 * kk;
      if (i < n1) {
        const float x = fdiv_z(ln[3 * i], f1), y = fdiv_z(ln[3 * i + 1], f1), z = fdiv_z(ln[3 * i + 2], f1);
        const float tx = (x * R[0] + y * R[1]) + z * R[2], ty = (x * R[4] + y * R[5]) + z * R[6],
                    tz = (x * R[8] + y * R[9]) + z * R[10];
        ss2 += dot3(tx, ty, tz, tx, ty, tz);
        keep[kk][0] = tx; keep[kk][1] = ty; keep[kk][2] = tz;
      }
    }
    for (int i = n0 + threadIdx.x + 256 * KEEP; i < n1; i += 256) {
      const float x = fdiv_z(ln[3 * i], f1), y = fdiv_z(ln[3 * i + 1], f1), z = fdiv_z(ln[3 * i + 2], f1);
      const float tx = (x * R[0] + y * R[1]) + z * R[2], ty = (x * R[4] + y * R[5]) + z * R[6],
                  tz = (x * R[8] + y * R[9]) + z * R[10];
      ss2 += dot3(tx, ty, tz, tx, ty, tz);
    }
    const float f2 = sqrtf(block_sum_256(ss2, red));
    if (out_scales) {
      if (threadIdx.x == 0) {
        out_scales[((long long)b * m.n_objects + o) * 2] = f1;
        out_scales[((long long)b * m.n_objects + o) * 2 + 1] = f2;
      }
      continue;
    }
    float* out = m.out_norms + (long long)b * m.n_norms * 3;
#pragma unroll
    for (int kk = 0; kk < KEEP; ++kk) {
      const int i = n0 + threadIdx.x + 256 * kk;
      if (i < n1) {
        out[3 * i] = fdiv_z(keep[kk][0], f2);
        out[3 * i + 1] = fdiv_z(keep[kk][1], f2);
        out[3 * i + 2] = fdiv_z(keep[kk][2], f2);
      }
    }
    for (int i = n0 + threadIdx.x + 256 * KEEP; i < n1; i += 256) {
      const float x = fdiv_z(ln[3 * i], f1), y = fdiv_z(ln[3 * i + 1], f1), z = fdiv_z(ln[3 * i + 2], f1);
      out[3 * i] = fdiv_z((x * R[0] + y * R[1]) + z * R[2], f2);
      out[3 * i + 1] = fdiv_z((x * R[4] + y * R[5]) + z * R[6], f2);
      out[3 * i + 2] = fdiv_z((x * R[8] + y * R[9]) + z * R[10], f2);
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

static bool g_no_attr_early() { static const bool v = getenv("JR_NO_ATTR") != nullptr; return v; }

static int check_common(const JrRenderArgs* a) {
  if (!a) return JR_ERR_NULL;
  if (a->shader < 0 || a->shader >= JR_NUM_SHADERS) return JR_ERR_SHADER;
  if (a->B <= 0 || a->W <= 0 || a->H <= 0 || a->T < 0 || a->n_pos < 0) return JR_ERR_DIMS;
  if (a->W > 32767 || a->H > 32767) return JR_ERR_DIMS;
  if (!a->world_to_clip.ptr || !a->viewport.ptr || !a->zbuffer) return JR_ERR_NULL;
  if (!a->tri_id && a->shader != JR_DEPTH) return JR_ERR_NULL;
  if (a->T > 0 && (!a->position.ptr || !a->faces.ptr)) return JR_ERR_NULL;
  if (a->T > 0 && a->n_pos <= 0) return JR_ERR_DIMS;
  const int s = a->shader;
  if (s != JR_DEPTH) {
    if ((!a->canvas && !a->canvas_u8) || !a->normal.ptr || !a->light_colour.ptr) return JR_ERR_NULL;
    if (a->canvas_u8 && (s == JR_PHONG_DARBOUX || g_no_attr_early() || a->T <= 0)) return JR_ERR_UNSUPPORTED;
    if (a->T > 0 && a->n_nrm <= 0) return JR_ERR_DIMS;
  }
  if (s == JR_GOURAUD && (!a->colour.ptr || !a->light_direction.ptr)) return JR_ERR_NULL;
  if (s >= JR_GOURAUD_TEXTURE) {
    if (!a->uv.ptr || !a->texture.ptr) return JR_ERR_NULL;
    if (a->tex_w <= 0 || a->tex_h <= 0 || (a->T > 0 && a->n_uv <= 0)) return JR_ERR_DIMS;
  }
  if ((s >= JR_GOURAUD_TEXTURE && s <= JR_PHONG_DARBOUX) && !a->light_direction.ptr) return JR_ERR_NULL;
  if (s >= JR_PHONG && !a->world_to_eye_norm.ptr) return JR_ERR_NULL;
  if (s == JR_PHONG_DARBOUX && (!a->normal_map.ptr || !a->id_to_face.ptr || !a->faces_indices.ptr))
    return JR_ERR_NULL;
  if (s >= JR_PHONG_REFLECTION) {
    if (!a->light_dir_eye.ptr || !a->ambient.ptr || !a->diffuse.ptr || !a->specular.ptr ||
        !a->specular_map.ptr || !a->texture_shape.ptr || !a->texture_index.ptr)
      return JR_ERR_NULL;
    if (a->spec_w <= 0 || a->spec_h <= 0 || a->n_objects <= 0 || a->n_texidx <= 0) return JR_ERR_DIMS;
  }
  if (s == JR_PHONG_REFLECTION_SHADOW) {
    if (!a->shadow_map.ptr || !a->shadow_strength.ptr || !a->shadow_world_to_clip.ptr ||
        !a->shadow_viewport.ptr)
      return JR_ERR_NULL;
    if (a->shadow_w <= 0 || a->shadow_h <= 0) return JR_ERR_DIMS;
  }
  if (a->inst_transform.ptr) {  // instanced geometry
    if (a->n_inst <= 0) return JR_ERR_DIMS;
    if (!a->inst_vert_object.ptr || !a->inst_scaling.ptr) return JR_ERR_NULL;
    if (s != JR_DEPTH && (!a->inst_norm_object.ptr || !a->inst_normal_matrix.ptr || !a->inst_norm_scale.ptr)) return JR_ERR_NULL;
    if (s == JR_PHONG_DARBOUX) return JR_ERR_UNSUPPORTED;
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

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device: set it once per device this process renders on
template <typename K>
static void smem_attr(K kernel, int bytes) { cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); }
static void set_kernel_attributes_once() {
  static std::once_flag once[64];
  int dev = 0;
  cudaGetDevice(&dev);
  std::call_once(once[dev & 63], [] {
    smem_attr(k_vis2<true, true, V2_K32_THREADS>, 200 * 1024);
    smem_attr(k_vis2<true, false, V2_THREADS>, 200 * 1024);
    smem_attr(k_vis2<false, false, V2_THREADS>, 200 * 1024);
    smem_attr(k_vis3<true, true, false, false>, 200 * 1024);
    smem_attr(k_vis3<true, false, false, false>, 200 * 1024);
    smem_attr(k_vis3<false, false, false, false>, 200 * 1024);
    smem_attr(k_vis3<true, true, false, true>, 200 * 1024);
    smem_attr(k_vis3<true, false, false, true>, 200 * 1024);
    smem_attr(k_vis3<false, false, false, true>, 200 * 1024);
    smem_attr(k_vis3<true, true, false, false, true>, 200 * 1024);
    smem_attr(k_vis3<true, true, false, true, true>, 200 * 1024);
    smem_attr(k_vis3<true, true, true, true>, 200 * 1024);
    smem_attr(k_vis3<true, false, true, true>, 200 * 1024);
    smem_attr(k_vis3<false, false, true, true>, 200 * 1024);
    smem_attr(k_raster_tile<true, true>, 64 * 1024);
    smem_attr(k_raster_tile<true, false>, 64 * 1024);
    smem_attr(k_raster_tile<false, false>, 64 * 1024);
  });
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

// debugging / A-B switches, read once at load
static const bool g_no_attr = getenv("JR_NO_ATTR") != nullptr;  // shade without attribute records
static const bool g_no_bins = getenv("JR_NO_BINS") != nullptr;  // every tile CTA scans all triangles
static const bool g_key64 = getenv("JR_KEY64") != nullptr;      // depth shader: keep packed 64-bit keys
static const bool g_vis2 = getenv("JR_VIS2") != nullptr;        // single-tile canvases: the one-phase kernel k_vis2
static const bool g_no_fused_mark = getenv("JR_NO_FUSED_MARK") != nullptr;  // visible-triangle lists by k_mark_visible
static const bool g_no_cluster = getenv("JR_NO_CLUSTER") != nullptr;        // small batches: one CTA per image as well

// Forward scratch: [binned visibility: triangle records + tile bitmasks][shading attribute records]
struct FwdLayout { size_t tiled, attr_off, flags_off, list_off, map_off, pc_off, total; bool use_attr, compact; int rec_stride; };
static FwdLayout fwd_layout(const JrRenderArgs* a) {
  FwdLayout F{};
  int tw, th, nx, ny;
  choose_tiles(a->W, a->H, &tw, &th, &nx, &ny);
  // binned path: records + bitmasks; single-tile path (k_vis3): the per-image spill list
  F.tiled = (nx * ny == 1) ? (g_vis2 ? 0 : v3_workspace_bytes(a->B, a->T)) : tiled_layout(a->B, a->W, a->H, a->T).total;
  // per-triangle attribute records pay off when a triangle is shared by several pixels
  // Per-triangle attribute records, built for the VISIBLE triangles only.  More pixels than triangles: the
  // record of triangle t sits in slot t.  More triangles than pixels (32x32 Brax frames): at most W*H
  // triangles are visible, records are packed in list order and a triangle -> slot map is kept.
  const long long npix = (long long)a->W * a->H;
  F.use_attr = a->shader != JR_DEPTH && a->shader != JR_PHONG_DARBOUX && a->T > 0 && !g_no_attr;
  F.compact = npix < a->T;
  F.rec_stride = F.compact ? (int)npix : a->T;
  F.attr_off = (F.tiled + 255) & ~(size_t)255;
  F.flags_off = F.attr_off + (((size_t)a->B * F.rec_stride * TA_FLOATS * 4 + 255) & ~(size_t)255);
  // [flag bits (B*T) | per-image counters (B)] zeroed per call, then the visible-triangle lists (B*T ints)
  const size_t flag_bytes = ((((size_t)a->B * a->T + 31) / 32) * 4 + (size_t)a->B * 4 + 255) & ~(size_t)255;
  F.list_off = F.flags_off + flag_bytes;
  F.map_off = F.list_off + (((size_t)a->B * a->T * 4 + 255) & ~(size_t)255);
  // per-image pixel-stage constants (PixConst), then the per-image pixel -> NDC tables (W + H floats), behind the slot map
  F.pc_off = (F.map_off + (F.compact ? (size_t)a->B * a->T * 4 : 0) + 255) & ~(size_t)255;
  F.total = F.use_attr ? F.pc_off + (size_t)a->B * (sizeof(PixConst) + (size_t)(a->W + a->H) * 4) : F.tiled;
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
  set_kernel_attributes_once();
  const bool depth = a->shader == JR_DEPTH;
  const char* vis_name = "k_vis3";
  jr::mark(stream, nullptr);
  bool fused_mark = false;   // the visibility kernel built the visible-triangle lists itself
  if (!depth) {
    // everything the shading stage will need is validated BEFORE the first launch: a failing call enqueues nothing
    const FwdLayout F0 = fwd_layout(a);
    if (F0.use_attr) {
      if (!a->workspace || a->workspace_bytes < F0.total) return JR_ERR_WORKSPACE;
      if (a->B > 65535) return JR_ERR_DIMS;
    }
  }
  if (nx * ny > 1 && !g_no_bins) {
    // two-level path: per-triangle records + per-tile bitmasks, then one CTA per (image, tile)
    const TiledLayout TLy = tiled_layout(a->B, a->W, a->H, a->T);
    if (!a->workspace || a->workspace_bytes < TLy.total) return JR_ERR_WORKSPACE;
    if (a->B > 65535) return JR_ERR_DIMS;
    char* ws = (char*)a->workspace;
    TriRecord* recs = (TriRecord*)(ws + TLy.rec);
    unsigned* masks = (unsigned*)(ws + TLy.mask);
    cudaMemsetAsync(masks, 0, (size_t)a->B * TLy.tiles * TLy.words * 4, stream);
    if (a->T > 0) {
      dim3 g1((a->T + 255) / 256, a->B);
      if (depth) k_setup_bin<true><<<g1, 256, 0, stream>>>(*a, recs, masks, TLy);
      else k_setup_bin<false><<<g1, 256, 0, stream>>>(*a, recs, masks, TLy);
      jr::g_launches++;
      jr::mark(stream, "memset+k_setup_bin");
    }
    vis_name = "k_raster_tile";
    const long long ctas2 = (long long)a->B * TLy.tiles;
    if (ctas2 > 2147483647LL) return JR_ERR_DIMS;
    const bool k32t = depth && !a->tri_id && !g_key64;
    const size_t sm = tl_smem(k32t ? 4 : 8).total;
    // shaders that go through attribute records: the resolve also builds the visible-triangle lists (TLVis)
    TLVis tvis{nullptr, nullptr, nullptr, nullptr};
    if (!depth) {
      const FwdLayout F = fwd_layout(a);
      if (F.use_attr && !g_no_fused_mark) {
        unsigned* flag_words = (unsigned*)(ws + F.flags_off);
        const size_t n_words = ((size_t)a->B * a->T + 31) / 32;
        cudaMemsetAsync(flag_words, 0, n_words * 4 + (size_t)a->B * 4, stream);
        tvis.flags = flag_words;
        tvis.count = (int*)(flag_words + n_words);
        tvis.list = (int*)(ws + F.list_off);
        tvis.slot_map = F.compact ? (int*)(ws + F.map_off) : nullptr;
        fused_mark = true;
      }
    }
    if (k32t) k_raster_tile<true, true><<<(unsigned)ctas2, TL_THREADS, sm, stream>>>(*a, recs, masks, TLy, tvis);
    else if (depth) k_raster_tile<true, false><<<(unsigned)ctas2, TL_THREADS, sm, stream>>>(*a, recs, masks, TLy, tvis);
    else k_raster_tile<false, false><<<(unsigned)ctas2, TL_THREADS, sm, stream>>>(*a, recs, masks, TLy, tvis);
  } else if (nx * ny == 1 && !g_vis2) {
    // single-tile canvases: filtered two-phase kernel (jr_vis3.cuh); z-only 32-bit keys for the depth shader
    // without a triangle-id output
    const bool k32 = depth && !a->tri_id && !g_key64;
    const V3Layout L = v3_layout(a->W, a->H, k32 ? 4 : 8);  // bytes: L.total
    if (a->T > 0 && (!a->workspace || a->workspace_bytes < v3_workspace_bytes(a->B, a->T))) return JR_ERR_WORKSPACE;
    const unsigned g = (unsigned)ctas;
    const bool inst = a->inst_transform.ptr != nullptr;
    // shaders that go through attribute records: the resolve also emits the visible-triangle lists (V3Vis)
    V3Vis vis{nullptr, nullptr, nullptr};
    if (!depth) {
      const FwdLayout F = fwd_layout(a);
      if (F.use_attr && a->T <= V3_VIS_MAXT && !g_no_fused_mark) {
        if (!a->workspace || a->workspace_bytes < F.total) return JR_ERR_WORKSPACE;
        const size_t n_words = ((size_t)a->B * a->T + 31) / 32;
        vis.count = (int*)((unsigned*)((char*)a->workspace + F.flags_off) + n_words);
        vis.list = (int*)((char*)a->workspace + F.list_off);
        vis.slot_map = F.compact ? (int*)((char*)a->workspace + F.map_off) : nullptr;
        fused_mark = true;
      }
    }
    // Batches that leave SMs idle: z-only-key depth passes split every image over a 2-CTA cluster
    // (k_vis3<..., CLUSTER>) while ALL half-image CTAs are still resident at once (2 B <= 4 per SM).  Measured on the
    // bench scene (84x84, 1932 triangles): B = 1..64 0.033 -> 0.027 ms, B = 192 0.039 -> 0.035, B = 256 equal; beyond
    // one resident wave the split LOSES (B = 512: 0.054 -> 0.067 ms: a half-image CTA takes ~0.62 of a whole one --
    // prologue, span table and barriers do not halve -- and two waves of those are slower than one wave of 512).
    int n_sm = 148;
    {
      static int sm_cache[64] = {0};
      int dev = 0;
      cudaGetDevice(&dev);
      if (!sm_cache[dev & 63]) cudaDeviceGetAttribute(&sm_cache[dev & 63], cudaDevAttrMultiProcessorCount, dev);
      if (sm_cache[dev & 63] > 0) n_sm = sm_cache[dev & 63];
    }
    if (k32 && !a->stats && !g_no_cluster && a->T > 0 && 2 * ctas <= (long long)n_sm * JR_V3_K32_CTAS) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2u * g, 1, 1);
      cfg.blockDim = dim3(V3_THREADS, 1, 1);
      cfg.dynamicSmemBytes = (size_t)L.total;
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      const cudaError_t e = inst ? cudaLaunchKernelEx(&cfg, k_vis3<true, true, false, true, true>, *a, vis)
                                 : cudaLaunchKernelEx(&cfg, k_vis3<true, true, false, false, true>, *a, vis);
      if (e != cudaSuccess) return JR_ERR_CUDA;
    } else if (a->stats) {  // counting variant (measurement aid): always the instancing-capable build
      if (k32) k_vis3<true, true, true, true><<<g, V3_THREADS, L.total, stream>>>(*a, vis);
      else if (depth) k_vis3<true, false, true, true><<<g, V3_THREADS, L.total, stream>>>(*a, vis);
      else k_vis3<false, false, true, true><<<g, V3_THREADS, L.total, stream>>>(*a, vis);
    } else if (inst) {
      if (k32) k_vis3<true, true, false, true><<<g, V3_THREADS, L.total, stream>>>(*a, vis);
      else if (depth) k_vis3<true, false, false, true><<<g, V3_THREADS, L.total, stream>>>(*a, vis);
      else k_vis3<false, false, false, true><<<g, V3_THREADS, L.total, stream>>>(*a, vis);
    } else {
      if (k32) k_vis3<true, true, false, false><<<g, V3_THREADS, L.total, stream>>>(*a, vis);
      else if (depth) k_vis3<true, false, false, false><<<g, V3_THREADS, L.total, stream>>>(*a, vis);
      else k_vis3<false, false, false, false><<<g, V3_THREADS, L.total, stream>>>(*a, vis);
    }
  } else {
    vis_name = "k_vis2";
    if (a->inst_transform.ptr) return JR_ERR_UNSUPPORTED;  // k_vis2 (A/B switch) reads merged arrays only
    if (a->depth_fill || a->depth_offset != 0.f) return JR_ERR_UNSUPPORTED;
    // depth shader without a triangle-id output: z-only 32-bit keys (half the shared memory, native atomic min)
    const bool k32 = depth && !a->tri_id && !g_key64;
    const V2Layout L = v2_layout(tw, th, k32 ? 4 : 8);
    if (k32) k_vis2<true, true, V2_K32_THREADS><<<(unsigned)ctas, V2_K32_THREADS, L.total, stream>>>(*a, tw, th, nx, ny);
    else if (depth) k_vis2<true, false, V2_THREADS><<<(unsigned)ctas, V2_THREADS, L.total, stream>>>(*a, tw, th, nx, ny);
    else k_vis2<false, false, V2_THREADS><<<(unsigned)ctas, V2_THREADS, L.total, stream>>>(*a, tw, th, nx, ny);
  }
  jr::g_launches++;
  jr::mark(stream, vis_name);
  if (!depth) {
    const int threads = 256;
    const int npix = a->W * a->H;
    int bx = (npix + threads - 1) / threads;
    if (bx > 4096) bx = 4096;
    const dim3 blocks(bx, a->B > 65535 ? 65535 : a->B);
    const FwdLayout F = fwd_layout(a);
    if (F.use_attr) {
      if (!a->workspace || a->workspace_bytes < F.total) return JR_ERR_WORKSPACE;
      if (a->B > 65535) return JR_ERR_DIMS;
      float* attrs = (float*)((char*)a->workspace + F.attr_off);
      unsigned* flag_words = (unsigned*)((char*)a->workspace + F.flags_off);
      const size_t n_words = ((size_t)a->B * a->T + 31) / 32;
      int* count = (int*)(flag_words + n_words);
      int* list = (int*)((char*)a->workspace + F.list_off);
      int* slot_map = F.compact ? (int*)((char*)a->workspace + F.map_off) : nullptr;
      if (!fused_mark) {
        cudaMemsetAsync(flag_words, 0, n_words * 4 + (size_t)a->B * 4, stream);
        k_mark_visible<0><<<dim3(bx > 64 ? 64 : bx, blocks.y), 256, 0, stream>>>(a->tri_id, flag_words, list, count, npix,
                                                                             a->T, a->B, slot_map);
        jr::g_launches++;
        jr::mark(stream, "memset+k_mark_visible");
      }
      const int rec_stride = F.rec_stride;
      const bool compact = F.compact;
      PixConst* pcs = (PixConst*)((char*)a->workspace + F.pc_off);
      const int g1x = (a->T + 127) / 128;
      dim3 g1(g1x < 8 ? g1x : 8, a->B);   // the kernel strides over the visible list
      const int tiles_x = (a->W + 31) / 32, tiles_y = (a->H + 31) / 32;
      const dim3 gu8(tiles_x * tiles_y, a->B);
      const dim3 grec((npix + threads - 1) / threads, a->B > 65535 ? 65535 : a->B, (a->B + 65534) / 65535);
#define JR_ATTR_CASE(S)                                                          \
  case S:                                                                        \
    k_tri_attr<S><<<g1, 128, 0, stream>>>(*a, attrs, list, count, rec_stride, compact, pcs); \
    jr::mark(stream, "k_tri_attr");                                              \
    if (a->canvas_u8) k_shade_rec_u8<S><<<gu8, 256, 0, stream>>>(*a, attrs, rec_stride, slot_map, tiles_x, tiles_y, pcs); \
    else k_shade_rec<S><<<grec, threads, 0, stream>>>(*a, attrs, rec_stride, slot_map, pcs); \
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
      jr::mark(stream, a->canvas_u8 ? "k_shade_rec_u8" : "k_shade_rec");
    } else {
      switch (a->shader) {
        case JR_GOURAUD: k_shade<JR_GOURAUD><<<blocks, threads, 0, stream>>>(*a); break;
        case JR_GOURAUD_TEXTURE: k_shade<JR_GOURAUD_TEXTURE><<<blocks, threads, 0, stream>>>(*a); break;
        case JR_PHONG: k_shade<JR_PHONG><<<blocks, threads, 0, stream>>>(*a); break;
        case JR_PHONG_DARBOUX: k_shade<JR_PHONG_DARBOUX><<<blocks, threads, 0, stream>>>(*a); break;
        case JR_PHONG_REFLECTION: k_shade<JR_PHONG_REFLECTION><<<blocks, threads, 0, stream>>>(*a); break;
        case JR_PHONG_REFLECTION_SHADOW:
          k_shade<JR_PHONG_REFLECTION_SHADOW><<<blocks, threads, 0, stream>>>(*a); break;
        default: return JR_ERR_SHADER;
      }
      jr::g_launches++;
      jr::mark(stream, "k_shade");
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

int jr_merge_objects(const JrMergeArgs* m, jr_stream_t stream_) {
  if (!m) return JR_ERR_NULL;
  if (m->B <= 0 || m->B > 65535 || m->n_objects <= 0 || m->n_verts < 0 || m->n_norms < 0) return JR_ERR_DIMS;
  if (!m->scaling.ptr || !m->transform.ptr || !m->normal_matrix.ptr) return JR_ERR_NULL;
  cudaStream_t stream = (cudaStream_t)stream_;
  jr::mark(stream, nullptr);
  if (m->n_verts > 0) {
    if (!m->local_verts.ptr || !m->vert_object.ptr || !m->out_verts) return JR_ERR_NULL;
    k_merge_verts<<<dim3((m->n_verts + 255) / 256, m->B), 256, 0, stream>>>(*m);
    jr::g_launches++;
    jr::mark(stream, "k_merge_verts");
  }
  if (m->n_norms > 0) {
    if (!m->local_norms.ptr || !m->norm_start.ptr || !m->out_norms) return JR_ERR_NULL;
    k_merge_norms<<<dim3(m->n_objects, (m->B + MN_GROUP - 1) / MN_GROUP), 256, 0, stream>>>(*m, nullptr);
    jr::g_launches++;
    jr::mark(stream, "k_merge_norms");
  }
  return cudaGetLastError() == cudaSuccess ? JR_OK : JR_ERR_CUDA;
}

int jr_instance_norm_scales(const JrMergeArgs* m, float* out_scales, jr_stream_t stream_) {
  if (!m || !out_scales) return JR_ERR_NULL;
  if (m->B <= 0 || m->B > 65535 * MN_GROUP || m->n_objects <= 0 || m->n_norms <= 0) return JR_ERR_DIMS;
  if (!m->local_norms.ptr || !m->norm_start.ptr || !m->normal_matrix.ptr) return JR_ERR_NULL;
  jr::mark((cudaStream_t)stream_, nullptr);
  k_merge_norms<<<dim3(m->n_objects, (m->B + MN_GROUP - 1) / MN_GROUP), 256, 0, (cudaStream_t)stream_>>>(*m, out_scales);
  jr::g_launches++;
  jr::mark((cudaStream_t)stream_, "k_merge_norms(scales)");
  return cudaGetLastError() == cudaSuccess ? JR_OK : JR_ERR_CUDA;
}

int jr_debug_audit_cull(const JrRenderArgs* a, unsigned long long* counters, jr_stream_t stream) {
  int st = check_common(a);
  if (st != JR_OK) return st;
  if (!counters) return JR_ERR_NULL;
  if (a->B > 65535) return JR_ERR_DIMS;
  if (a->T > 0) {
    k_audit_cull<<<dim3((a->T + 7) / 8, a->B), 256, 0, (cudaStream_t)stream>>>(*a, counters);
    jr::g_launches++;
  }
  return cudaGetLastError() == cudaSuccess ? JR_OK : JR_ERR_CUDA;
}

int jr_add_scalar(float* data, long long n, float value, jr_stream_t stream) {
  if (!data) return JR_ERR_NULL;
  if (n <= 0) return JR_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  jr::mark((cudaStream_t)stream, nullptr);
  k_add_scalar<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(data, n, value);
  jr::g_launches++;
  jr::mark((cudaStream_t)stream, "k_add_scalar");
  return cudaGetLastError() == cudaSuccess ? JR_OK : JR_ERR_CUDA;
}

int jr_canvas_to_uint8_display(const float* canvas, uint8_t* out, int B, int W, int H, jr_stream_t stream) {
  if (!canvas || !out) return JR_ERR_NULL;
  if (B <= 0 || W <= 0 || H <= 0 || B > 65535) return JR_ERR_DIMS;
  dim3 grid((W + 31) / 32, (H + 31) / 32, B), block(32, 8);
  jr::mark((cudaStream_t)stream, nullptr);
  k_to_uint8_display<<<grid, block, 0, (cudaStream_t)stream>>>(canvas, out, B, W, H);
  jr::g_launches++;
  jr::mark((cudaStream_t)stream, "k_to_uint8_display");
  return cudaGetLastError() == cudaSuccess ? JR_OK : JR_ERR_CUDA;
}

}  // extern "C"
