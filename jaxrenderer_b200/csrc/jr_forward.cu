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
  // large triangles: whole CTA, lanes over pixels (y fastest)
  const int nbig = min(bigq_n, BIGQ_CAP);
  for (int e = 0; e < nbig; ++e) {
    const BigTri& q = bigq[e];
    const int bh = q.y1 - q.y0 + 1;
    const int n = (q.x1 - q.x0 + 1) * bh;
    for (int i = tid; i < n; i += VIS_THREADS) {
      const int lx = i / bh;
      const int x = q.x0 + lx, y = q.y0 + (i - lx * bh);
      const float xn = xs[x];
      raster_pixel(q.inv, q.zc, xn * q.inv[0], xn * q.inv[1], xn * q.inv[2], ys[y], vp22, vp23, q.tri,
                   &keys[x * tile_h + y]);
    }
  }
  __syncthreads();
  // resolve
  int32_t* __restrict__ tri_out = a.tri_id ? a.tri_id + (long long)b * a.W * a.H : nullptr;
  float* __restrict__ z_out = DEPTH ? a.zbuffer + (long long)b * a.W * a.H : nullptr;
  const bool use0 = DEPTH && tri0_flag;
  for (int i = tid; i < tw * th; i += VIS_THREADS) {
    const int lx = i / th, ly = i - lx * th;
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

// --------------------------------------------------------------------- shading
struct Light3 { float v[3]; };
__device__ __forceinline__ void load3(const JrF32& arr, int b, float out[3]) {
  const float* p = arr.ptr + (long long)b * arr.batch_stride;
  out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
}
__device__ __forceinline__ Vec3 loadv3(const float* p, int i) {
  return Vec3{p[3 * i], p[3 * i + 1], p[3 * i + 2]};
}

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

    const float* __restrict__ w2c = a.world_to_clip.ptr + (long long)b * a.world_to_clip.batch_stride;
    const float* __restrict__ vp = a.viewport.ptr + (long long)b * a.viewport.batch_stride;
    const float* __restrict__ pos = a.position.ptr + (long long)b * a.position.batch_stride;
    const int32_t* __restrict__ faces = a.faces.ptr + (long long)b * a.faces.batch_stride;
    const int fi[3] = {faces[3 * tri], faces[3 * tri + 1], faces[3 * tri + 2]};
    float cl[3][4];
    Vec3 P[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      P[k] = loadv3(pos, fi[k]);
      to_clip(w2c, P[k].x, P[k].y, P[k].z, cl[k]);
    }
    float M[9], inv[9];
    tri_matrix(cl[0], cl[1], cl[2], M);
    lu_inverse3(M, inv);
    const float xn = ((float)x - vp[3]) / vp[0];
    const float yn = ((float)y - vp[7]) / vp[5];
    float cc[3];
    clip_coef(inv, xn, yn, cc);
    const float w_rec = (cc[0] + cc[1]) + cc[2];
    const float z = (cc[0] * cl[0][2] + cc[1] * cl[1][2]) + cc[2] * cl[2][2];
    const float zw = z * vp[10] + vp[11];
    const float tc[3] = {cc[0] / w_rec, cc[1] / w_rec, cc[2] / w_rec};

    float col[3] = {0.f, 0.f, 0.f};
    bool keep = true;

    // index rows for the other attributes (NULL -> faces)
    int fn[3] = {fi[0], fi[1], fi[2]}, fu[3] = {fi[0], fi[1], fi[2]};
    if (SHADER != JR_DEPTH) {
      if (a.faces_norm.ptr) {
        const int32_t* f = a.faces_norm.ptr + (long long)b * a.faces_norm.batch_stride + 3 * tri;
        fn[0] = f[0]; fn[1] = f[1]; fn[2] = f[2];
      }
      if (a.faces_uv.ptr) {
        const int32_t* f = a.faces_uv.ptr + (long long)b * a.faces_uv.batch_stride + 3 * tri;
        fu[0] = f[0]; fu[1] = f[1]; fu[2] = f[2];
      }
    }

    if (SHADER == JR_GOURAUD || SHADER == JR_GOURAUD_TEXTURE) {
      float ld[3], lcol[3];
      load3(a.light_direction, b, ld);
      load3(a.light_colour, b, lcol);
      const Vec3 nl = normalise3(Vec3{ld[0], ld[1], ld[2]});
      const float* __restrict__ nrm = a.normal.ptr + (long long)b * a.normal.batch_stride;
      float inten[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const Vec3 n = normalise3(loadv3(nrm, fn[k]));
        inten[k] = dot3(n.x, n.y, n.z, nl.x, nl.y, nl.z);
      }
      if (SHADER == JR_GOURAUD) {
        const float* __restrict__ cv = a.colour.ptr + (long long)b * a.colour.batch_stride;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float v0 = (cv[3 * fi[0] + c] * lcol[c]) * inten[0];
          const float v1 = (cv[3 * fi[1] + c] * lcol[c]) * inten[1];
          const float v2 = (cv[3 * fi[2] + c] * lcol[c]) * inten[2];
          col[c] = interp3(tc, v0, v1, v2);
          keep = keep && (col[c] >= 0.f);
        }
      } else {
        const float* __restrict__ uvp = a.uv.ptr + (long long)b * a.uv.batch_stride;
        const float* __restrict__ tex = a.texture.ptr + (long long)b * a.texture.batch_stride;
        const float u = interp3(tc, uvp[2 * fu[0]], uvp[2 * fu[1]], uvp[2 * fu[2]]);
        const float v = interp3(tc, uvp[2 * fu[0] + 1], uvp[2 * fu[1] + 1], uvp[2 * fu[2] + 1]);
        const int ui = pymod((int)floorf(u), a.tex_w), vi = pymod((int)floorf(v), a.tex_h);
        const float* texel = tex + ((long long)ui * a.tex_h + vi) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float lc = interp3(tc, lcol[c] * inten[0], lcol[c] * inten[1], lcol[c] * inten[2]);
          keep = keep && (lc >= 0.f);
          col[c] = texel[c] * lc;
        }
      }
    } else if (SHADER >= JR_PHONG) {
      const float* __restrict__ wen = a.world_to_eye_norm.ptr + (long long)b * a.world_to_eye_norm.batch_stride;
      const float* __restrict__ nrm = a.normal.ptr + (long long)b * a.normal.batch_stride;
      const float* __restrict__ uvp = a.uv.ptr + (long long)b * a.uv.batch_stride;
      const float* __restrict__ tex = a.texture.ptr + (long long)b * a.texture.batch_stride;
      Vec3 nv[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) nv[k] = apply_vec(wen, normalise3(loadv3(nrm, fn[k])));
      const Vec3 normal = {interp3(tc, nv[0].x, nv[1].x, nv[2].x), interp3(tc, nv[0].y, nv[1].y, nv[2].y),
                           interp3(tc, nv[0].z, nv[1].z, nv[2].z)};
      const float u = interp3(tc, uvp[2 * fu[0]], uvp[2 * fu[1]], uvp[2 * fu[2]]);
      const float v = interp3(tc, uvp[2 * fu[0] + 1], uvp[2 * fu[1] + 1], uvp[2 * fu[2] + 1]);
      Vec3 nn = normalise3(normal);
      float lcol[3];
      load3(a.light_colour, b, lcol);

      if (SHADER == JR_PHONG || SHADER == JR_PHONG_DARBOUX) {
        float ld[3];
        load3(a.light_direction, b, ld);
        const Vec3 nl = normalise3(Vec3{ld[0], ld[1], ld[2]});
        const int ui = pymod((int)floorf(u), a.tex_w), vi = pymod((int)floorf(v), a.tex_h);
        if (SHADER == JR_PHONG_DARBOUX) {
          // phong_darboux.py:144-151, :231-262
          const int32_t* __restrict__ i2f = a.id_to_face.ptr + (long long)b * a.id_to_face.batch_stride;
          const int32_t* __restrict__ fidx = a.faces_indices.ptr + (long long)b * a.faces_indices.batch_stride;
          const int face = i2f[fi[0]];
          float tr[3][3], tuv[3][2];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int vtx = fidx[3 * face + k];
            float tcq[4];
            to_clip(w2c, pos[3 * vtx], pos[3 * vtx + 1], pos[3 * vtx + 2], tcq);
            const bool w0 = tcq[3] == 0.0f;
            tr[k][0] = w0 ? tcq[0] : tcq[0] / tcq[3];
            tr[k][1] = w0 ? tcq[1] : tcq[1] / tcq[3];
            tr[k][2] = w0 ? tcq[2] : tcq[2] / tcq[3];
            tuv[k][0] = uvp[2 * vtx]; tuv[k][1] = uvp[2 * vtx + 1];
          }
          const float A[9] = {tr[1][0] - tr[0][0], tr[1][1] - tr[0][1], tr[1][2] - tr[0][2],
                              tr[2][0] - tr[0][0], tr[2][1] - tr[0][1], tr[2][2] - tr[0][2],
                              nn.x, nn.y, nn.z};
          float AI[9];
          lu_inverse3(A, AI);
          const float du0 = tuv[1][0] - tuv[0][0], du1 = tuv[2][0] - tuv[0][0];
          const float dv0 = tuv[1][1] - tuv[0][1], dv1 = tuv[2][1] - tuv[0][1];
          const Vec3 iv = normalise3(Vec3{AI[0] * du0 + AI[1] * du1, AI[3] * du0 + AI[4] * du1,
                                          AI[6] * du0 + AI[7] * du1});
          const Vec3 jv = normalise3(Vec3{AI[0] * dv0 + AI[1] * dv1, AI[3] * dv0 + AI[4] * dv1,
                                          AI[6] * dv0 + AI[7] * dv1});
          const float* nm = a.normal_map.ptr + (long long)b * a.normal_map.batch_stride +
                            ((long long)ui * a.tex_h + vi) * 3;
          const Vec3 bn = {(iv.x * nm[0] + jv.x * nm[1]) + nn.x * nm[2],
                           (iv.y * nm[0] + jv.y * nm[1]) + nn.y * nm[2],
                           (iv.z * nm[0] + jv.z * nm[1]) + nn.z * nm[2]};
          nn = normalise3(bn);
        }
        const float ndl = dot3(nn.x, nn.y, nn.z, nl.x, nl.y, nl.z);
        const float* texel = tex + ((long long)ui * a.tex_h + vi) * 3;
        float lc[3];
        bool ok = true;
#pragma unroll
        for (int c = 0; c < 3; ++c) { lc[c] = lcol[c] * ndl; ok = ok && (lc[c] >= 0.f); }
#pragma unroll
        for (int c = 0; c < 3; ++c) col[c] = ok ? texel[c] * lc[c] : 0.f;
      } else {
        // phong_reflection.py:175-220, phong_reflection_shadow.py:196-257
        const int32_t* __restrict__ ftp =
            a.faces_tex.ptr ? a.faces_tex.ptr + (long long)b * a.faces_tex.batch_stride + 3 * tri : nullptr;
        const int tv = ftp ? ftp[0] : fi[0];
        const int ti = (a.texture_index.ptr + (long long)b * a.texture_index.batch_stride)[tv];
        const int32_t* tsh = a.texture_shape.ptr + (long long)b * a.texture_shape.batch_stride + 2 * ti;
        float fu0 = u - truncf(u), fv0 = v - truncf(v);
        if (fu0 < 0.f) fu0 = fu0 + 1.f;
        if (fv0 < 0.f) fv0 = fv0 + 1.f;
        const float ur = fu0 * (float)tsh[0] + (float)(ti * a.texture_offset);
        const float vr = fv0 * (float)tsh[1];
        const int U = (int)floorf(ur), V = (int)floorf(vr);
        const float* texel = tex + ((long long)wrap_clamp(U, a.tex_w) * a.tex_h + wrap_clamp(V, a.tex_h)) * 3;
        float lde[3], amb[3], dif[3], spe[3];
        load3(a.light_dir_eye, b, lde);
        load3(a.ambient, b, amb);
        load3(a.diffuse, b, dif);
        load3(a.specular, b, spe);
        const Vec3 ld = normalise3(Vec3{lde[0], lde[1], lde[2]});
        const float ndl = dot3(nn.x, nn.y, nn.z, ld.x, ld.y, ld.z);
        const float diffuse = fmaxf(ndl, 0.f);
        const float two_ndl = 2.f * ndl;
        const Vec3 refl = normalise3(Vec3{two_ndl * nn.x - ld.x, two_ndl * nn.y - ld.y, two_ndl * nn.z - ld.z});
        const float sexp = (a.specular_map.ptr + (long long)b * a.specular_map.batch_stride)
            [(long long)wrap_clamp(U, a.spec_w) * a.spec_h + wrap_clamp(V, a.spec_h)];
        const float specular = powf(fmaxf(refl.z, 0.f), sexp);
        float shadow[3] = {1.f, 1.f, 1.f};
        if (SHADER == JR_PHONG_REFLECTION_SHADOW) {
          const float* __restrict__ sw2c = a.shadow_world_to_clip.ptr + (long long)b * a.shadow_world_to_clip.batch_stride;
          const float* __restrict__ svp = a.shadow_viewport.ptr + (long long)b * a.shadow_viewport.batch_stride;
          float scv[3][4];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            float s[4];
            to_clip(sw2c, P[k].x, P[k].y, P[k].z, s);
#pragma unroll
            for (int j = 0; j < 4; ++j) scv[k][j] = s[j] / s[3];
          }
          float sc[4], ss[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) sc[j] = interp3(tc, scv[0][j], scv[1][j], scv[2][j]);
#pragma unroll
          for (int r = 0; r < 4; ++r)
            ss[r] = ((svp[4 * r] * sc[0] + svp[4 * r + 1] * sc[1]) + svp[4 * r + 2] * sc[2]) + svp[4 * r + 3] * sc[3];
          const float sx = ss[0] / ss[3], sy = ss[1] / ss[3], sz = ss[2] / ss[3];
          // Shadow.get (shadow.py:129-153)
          float rx = roundf(sx), ry = roundf(sy);
          rx = fminf(fmaxf(rx, -1e9f), 1e9f);
          ry = fminf(fmaxf(ry, -1e9f), 1e9f);
          int px = (int)rx, py = (int)ry;
          if (px < 0) px += a.shadow_w;
          if (py < 0) py += a.shadow_h;
          float sval = __int_as_float(0x7f800000);
          if (px >= 0 && px < a.shadow_w && py >= 0 && py < a.shadow_h && rx == rx && ry == ry)
            sval = (a.shadow_map.ptr + (long long)b * a.shadow_map.batch_stride)[(long long)px * a.shadow_h + py];
          const bool lit = sz <= sval;
          float str[3];
          load3(a.shadow_strength, b, str);
#pragma unroll
          for (int c = 0; c < 3; ++c) shadow[c] = lit ? 1.f : 1.f - str[c];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float ds = dif[c] * diffuse + spe[c] * specular;
          if (SHADER == JR_PHONG_REFLECTION)
            col[c] = amb[c] * texel[c] + (ds * lcol[c]) * texel[c];
          else
            col[c] = amb[c] * texel[c] + ((shadow[c] * ds) * texel[c]) * lcol[c];
        }
      }
    }

    if (keep) {
      a.zbuffer[gi] = zw;
      float* o = a.canvas + gi * 3;
      o[0] = col[0]; o[1] = col[1]; o[2] = col[2];
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
  if ((size_t)W * H * 8 <= 96 * 1024) { *tw = W; *th = H; *nx = 1; *ny = 1; return; }
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

size_t jr_workspace_bytes(const JrRenderArgs* a) { (void)a; return 0; }

long long jr_launch_count(void) { return jr::g_launches.load(); }

int jr_render_forward(const JrRenderArgs* a, jr_stream_t stream_) {
  int st = check_common(a);
  if (st != JR_OK) return st;
  cudaStream_t stream = (cudaStream_t)stream_;
  int tw, th, nx, ny;
  choose_tiles(a->W, a->H, &tw, &th, &nx, &ny);
  const VisSmemLayout L = vis_layout(tw, th);
  const long long ctas = (long long)a->B * nx * ny;
  if (ctas > 2147483647LL) return JR_ERR_DIMS;
  if (a->shader == JR_DEPTH) {
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(k_visibility<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(k_visibility<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr_done = true;
    }
    k_visibility<true><<<(unsigned)ctas, VIS_THREADS, L.total, stream>>>(*a, tw, th, nx, ny);
    jr::g_launches++;
  } else {
    static bool attr_done2 = false;
    if (!attr_done2) {
      cudaFuncSetAttribute(k_visibility<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(k_visibility<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr_done2 = true;
    }
    k_visibility<false><<<(unsigned)ctas, VIS_THREADS, L.total, stream>>>(*a, tw, th, nx, ny);
    jr::g_launches++;
    const long long total = (long long)a->B * a->W * a->H;
    const int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    if (blocks > 148LL * 64) blocks = 148LL * 64;
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
