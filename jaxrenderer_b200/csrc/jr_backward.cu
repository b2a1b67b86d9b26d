// jr_backward.cu -- reverse mode of pipeline.render through FIXED visibility
// (the reference gets it from jax.grad over renderer/pipeline.py:470-537; the
// integer argmin / floor / round / comparisons cut the graph).
//
// Inputs: the saved triangle-id G-buffer + output cotangents.  Every pixel's
// fragment is recomputed with the forward's own code (jr_shade.cuh) and
// back-propagated analytically.  Accumulation is DETERMINISTIC, no float
// atomics anywhere:
//   * scene-global parameters (camera matrices, light, shadow strength):
//     per-thread fixed-order partial sums -> fixed-shape block tree reduction
//     -> per-block partials -> fixed-order final reduction;
//   * keyed targets (diffuse texture, specular map, vertex positions /
//     colours / normals): (key, pixel) pairs -> CUB radix sort (stable) ->
//     fused "recompute + segmented reduction" kernel over the sorted order
//     (fixed 256-entry chunks, Hillis-Steele segmented scan, fixed-order carry
//     resolution across chunks).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <cub/device/device_radix_sort.cuh>

#include "../../include/jr_b200.h"
#include "jr_common.cuh"
#include "jr_device.cuh"
#include "jr_shade.cuh"

namespace jr {

// ---- packed layout of the scene-global gradient vector
enum : int {
  G_W2C = 0,    // 16
  G_VP00 = 16, G_VP03 = 17, G_VP11 = 18, G_VP13 = 19, G_VP22 = 20, G_VP23 = 21,
  G_WEN = 22,   // 9: upper-left 3x3, row-major
  G_LDIR = 31, G_LCOL = 34, G_LDE = 37, G_AMB = 40, G_DIF = 43, G_SPE = 46, G_STR = 49,
  NG = 52
};

enum : int { MODE_TEXEL = 0, MODE_SPEC = 1, MODE_POS = 2, MODE_NRM = 3,
              MODE_NMAP = 4, MODE_UV = 5, MODE_POS2 = 6 };  // phong_darboux: normal map, uv and tangent-triangle positions

// Scene-global accumulators of one thread: a column of a [NG][BWD_THREADS] shared-memory array
// (conflict-free; keeps 52 values out of the register file, which is what limits occupancy here).
struct GAcc {
  float* p;
  __device__ __forceinline__ float& operator[](int j) const { return p[j * 256]; }
};

struct PixGrad {
  GAcc g;
  float d_tex[3];
  float d_sexp;
  float d_pos[3][3];
  float d_col[3][3];
  float d_nrm[3][3];
  // phong_darboux
  float d_nmap[3];
  float d_uv[3][2];
  float d_pos2[3][3];
};

// 1 / |v| for the reverse pass (continuous arithmetic: reciprocal square root + one Newton step, < 1 ulp)
__device__ __forceinline__ float inv_norm3(Vec3 v) {
  const float d = fmaf(v.x, v.x, fmaf(v.y, v.y, v.z * v.z));
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return r * fmaf(-0.5f * d, r * r, 1.5f);
}
__device__ __forceinline__ Vec3 normalise_bwd(Vec3 v, Vec3 dy) {
  const float inv_n = inv_norm3(v);
  const Vec3 y = {v.x * inv_n, v.y * inv_n, v.z * inv_n};
  const float d = y.x * dy.x + y.y * dy.y + y.z * dy.z;
  return Vec3{(dy.x - y.x * d) * inv_n, (dy.y - y.y * d) * inv_n, (dy.z - y.z * d) * inv_n};
}

// Back-propagate one pixel.  WG/WT/WV select which outputs are produced.
template <int S, bool WG, bool WT, bool WV>
__device__ __forceinline__ void backprop_pixel(const JrRenderArgs& a, int b, const Frag& f, float d_zw,
                                               const float d_col[3], PixGrad& o) {
  float d_tc[3] = {0.f, 0.f, 0.f};
  const float* tc = f.tc;
  if (WT) { o.d_tex[0] = o.d_tex[1] = o.d_tex[2] = 0.f; o.d_sexp = 0.f; }
  if (WV) {
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) { o.d_pos[k][c] = 0.f; o.d_col[k][c] = 0.f; o.d_nrm[k][c] = 0.f; }
  }

  if (S == JR_GOURAUD || S == JR_GOURAUD_TEXTURE) {
    float d_I[3] = {0.f, 0.f, 0.f};
    if (S == JR_GOURAUD) {
      const float* __restrict__ cv = a.colour.ptr + (long long)b * a.colour.batch_stride;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float ck = cv[3 * f.fi[k] + c];
          const float d_colv = tc[k] * d_col[c];
          d_tc[k] += d_col[c] * ((ck * f.lcol[c]) * f.inten[k]);
          if (WG) o.g[G_LCOL + c] += d_colv * ck * f.inten[k];
          d_I[k] += d_colv * ck * f.lcol[c];
          if (WV) o.d_col[k][c] = d_colv * f.lcol[c] * f.inten[k];
        }
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (WT) o.d_tex[c] = d_col[c] * f.lc[c];
        const float d_lc = d_col[c] * f.tex[c];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          d_tc[k] += d_lc * (f.lcol[c] * f.inten[k]);
          if (WG) o.g[G_LCOL + c] += tc[k] * d_lc * f.inten[k];
          d_I[k] += tc[k] * d_lc * f.lcol[c];
        }
      }
    }
    Vec3 d_nl = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      d_nl.x += d_I[k] * f.nvert[k].x; d_nl.y += d_I[k] * f.nvert[k].y; d_nl.z += d_I[k] * f.nvert[k].z;
      if (WV) {
        const Vec3 dn = normalise_bwd(f.nraw[k], Vec3{d_I[k] * f.nl.x, d_I[k] * f.nl.y, d_I[k] * f.nl.z});
        o.d_nrm[k][0] = dn.x; o.d_nrm[k][1] = dn.y; o.d_nrm[k][2] = dn.z;
      }
    }
    if (WG) {
      const Vec3 dl = normalise_bwd(Vec3{f.lraw[0], f.lraw[1], f.lraw[2]}, d_nl);
      o.g[G_LDIR] += dl.x; o.g[G_LDIR + 1] += dl.y; o.g[G_LDIR + 2] += dl.z;
    }
  } else if (S >= JR_PHONG) {
    Vec3 d_nn = {0.f, 0.f, 0.f};
    bool active = true;
    if (S == JR_PHONG_DARBOUX) {
      // phong_darboux.py:231-262 in reverse: normal = normalise(B @ nm), B = [i | j | n],
      // i, j = normalise(AI[:, :2] @ (du | dv)), AI = inv([tr1 - tr0; tr2 - tr0; n])
      active = f.ok;
      if (WT) { o.d_nmap[0] = o.d_nmap[1] = o.d_nmap[2] = 0.f; }
      if (WV) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          o.d_uv[k][0] = o.d_uv[k][1] = 0.f;
          o.d_pos2[k][0] = o.d_pos2[k][1] = o.d_pos2[k][2] = 0.f;
        }
      }
      if (active) {
        float d_ndl = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (WT) o.d_tex[c] = d_col[c] * f.lc[c];
          const float d_lc = d_col[c] * f.tex[c];
          if (WG) o.g[G_LCOL + c] += d_lc * f.ndl;
          d_ndl += d_lc * f.lcol[c];
        }
        const Vec3 n2 = f.dn2;
        if (WG) {
          const Vec3 dl = normalise_bwd(Vec3{f.lraw[0], f.lraw[1], f.lraw[2]},
                                        Vec3{d_ndl * n2.x, d_ndl * n2.y, d_ndl * n2.z});
          o.g[G_LDIR] += dl.x; o.g[G_LDIR + 1] += dl.y; o.g[G_LDIR + 2] += dl.z;
        }
        const Vec3 d_bn = normalise_bwd(f.dbn, Vec3{d_ndl * f.nl.x, d_ndl * f.nl.y, d_ndl * f.nl.z});
        const Vec3 iv = f.div_, jv = f.djv, nn = f.nn;
        if (WT) {
          o.d_nmap[0] = iv.x * d_bn.x + iv.y * d_bn.y + iv.z * d_bn.z;
          o.d_nmap[1] = jv.x * d_bn.x + jv.y * d_bn.y + jv.z * d_bn.z;
          o.d_nmap[2] = nn.x * d_bn.x + nn.y * d_bn.y + nn.z * d_bn.z;
        }
        d_nn = Vec3{f.dnm[2] * d_bn.x, f.dnm[2] * d_bn.y, f.dnm[2] * d_bn.z};
        const Vec3 d_ivr = normalise_bwd(f.divr, Vec3{f.dnm[0] * d_bn.x, f.dnm[0] * d_bn.y, f.dnm[0] * d_bn.z});
        const Vec3 d_jvr = normalise_bwd(f.djvr, Vec3{f.dnm[1] * d_bn.x, f.dnm[1] * d_bn.y, f.dnm[1] * d_bn.z});
        const float di[3] = {d_ivr.x, d_ivr.y, d_ivr.z}, dj[3] = {d_jvr.x, d_jvr.y, d_jvr.z};
        const float du0 = f.dduv[0], du1 = f.dduv[1], dv0 = f.dduv[2], dv1 = f.dduv[3];
        const float* AI = f.dAI;
        float dAI[9];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          dAI[3 * r] = di[r] * du0 + dj[r] * dv0;
          dAI[3 * r + 1] = di[r] * du1 + dj[r] * dv1;
          dAI[3 * r + 2] = 0.f;
        }
        if (WV) {
          const float d_du0 = di[0] * AI[0] + di[1] * AI[3] + di[2] * AI[6];
          const float d_du1 = di[0] * AI[1] + di[1] * AI[4] + di[2] * AI[7];
          const float d_dv0 = dj[0] * AI[0] + dj[1] * AI[3] + dj[2] * AI[6];
          const float d_dv1 = dj[0] * AI[1] + dj[1] * AI[4] + dj[2] * AI[7];
          o.d_uv[1][0] = d_du0; o.d_uv[1][1] = d_dv0;
          o.d_uv[2][0] = d_du1; o.d_uv[2][1] = d_dv1;
          o.d_uv[0][0] = -(d_du0 + d_du1); o.d_uv[0][1] = -(d_dv0 + d_dv1);
        }
        // d_A = -(AI^T dAI AI^T)
        float T1[9], dA[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int l = 0; l < 3; ++l) T1[3 * i + l] = AI[i] * dAI[l] + AI[3 + i] * dAI[3 + l] + AI[6 + i] * dAI[6 + l];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j)
            dA[3 * i + j] = -(T1[3 * i] * AI[3 * j] + T1[3 * i + 1] * AI[3 * j + 1] + T1[3 * i + 2] * AI[3 * j + 2]);
        d_nn.x += dA[6]; d_nn.y += dA[7]; d_nn.z += dA[8];
        if (WG || WV) {
          const float* __restrict__ w2c = a.world_to_clip.ptr + (long long)b * a.world_to_clip.batch_stride;
          float d_tr[3][3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            d_tr[1][c] = dA[c]; d_tr[2][c] = dA[3 + c]; d_tr[0][c] = -(dA[c] + dA[3 + c]);
          }
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float w = f.dtw[k];
            float dc[4];
            if (w != 0.0f) {
              dc[0] = d_tr[k][0] / w; dc[1] = d_tr[k][1] / w; dc[2] = d_tr[k][2] / w;
              dc[3] = -(d_tr[k][0] * f.dtr[k][0] + d_tr[k][1] * f.dtr[k][1] + d_tr[k][2] * f.dtr[k][2]) / w;
            } else {
              dc[0] = d_tr[k][0]; dc[1] = d_tr[k][1]; dc[2] = d_tr[k][2]; dc[3] = 0.f;
            }
            const float ph[4] = {f.dP[k].x, f.dP[k].y, f.dP[k].z, 1.f};
            if (WG) {
#pragma unroll
              for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) o.g[G_W2C + 4 * r + c] += dc[r] * ph[c];
            }
            if (WV) {
#pragma unroll
              for (int c = 0; c < 3; ++c)
                o.d_pos2[k][c] = w2c[c] * dc[0] + w2c[4 + c] * dc[1] + w2c[8 + c] * dc[2] + w2c[12 + c] * dc[3];
            }
          }
        }
      }
    } else if (S == JR_PHONG) {
      active = f.ok;
      if (active) {
        float d_ndl = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (WT) o.d_tex[c] = d_col[c] * f.lc[c];
          const float d_lc = d_col[c] * f.tex[c];
          if (WG) o.g[G_LCOL + c] += d_lc * f.ndl;
          d_ndl += d_lc * f.lcol[c];
        }
        d_nn = Vec3{d_ndl * f.nl.x, d_ndl * f.nl.y, d_ndl * f.nl.z};
        if (WG) {
          const Vec3 dl = normalise_bwd(Vec3{f.lraw[0], f.lraw[1], f.lraw[2]},
                                        Vec3{d_ndl * f.nn.x, d_ndl * f.nn.y, d_ndl * f.nn.z});
          o.g[G_LDIR] += dl.x; o.g[G_LDIR + 1] += dl.y; o.g[G_LDIR + 2] += dl.z;
        }
      }
    } else {  // S6 / S7
      float d_diffuse = 0.f, d_s = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float sh = (S == JR_PHONG_REFLECTION_SHADOW) ? f.shadow[c] : 1.f;
        const float tl = f.tex[c] * f.lcol[c];
        if (WG) {
          o.g[G_AMB + c] += d_col[c] * f.tex[c];
          o.g[G_LCOL + c] += d_col[c] * sh * f.ds[c] * f.tex[c];
          if (S == JR_PHONG_REFLECTION_SHADOW && !f.lit) o.g[G_STR + c] += -d_col[c] * f.ds[c] * tl;
        }
        if (WT) o.d_tex[c] = d_col[c] * (f.amb[c] + sh * f.ds[c] * f.lcol[c]);
        const float d_ds = d_col[c] * sh * tl;
        if (WG) { o.g[G_DIF + c] += d_ds * f.diffuse; o.g[G_SPE + c] += d_ds * f.specular; }
        d_diffuse += d_ds * f.dif[c];
        d_s += d_ds * f.spe[c];
      }
      // s = pow(base, e)
      float d_base = 0.f;
      // base^(e - 1) = base^e / base: the forward's power is at hand (a second powf only where base is 0)
      if (f.sexp != 0.f) d_base = d_s * f.sexp * (f.base > 0.f ? f.specular * (1.0f / f.base) : powf(f.base, f.sexp - 1.f));
      if (WT) o.d_sexp = (f.base > 0.f) ? d_s * f.specular * logf(f.base) : 0.f;
      const Vec3 rv = f.rv;
      const float refl_z = rv.z * inv_norm3(rv);
      const float d_refl_z = (refl_z > 0.f) ? d_base : 0.f;
      const Vec3 d_rv = normalise_bwd(rv, Vec3{0.f, 0.f, d_refl_z});
      const Vec3 nn = f.nn, ld = f.nl;
      const float d_ndl = ((f.ndl > 0.f) ? d_diffuse : 0.f) + 2.f * (d_rv.x * nn.x + d_rv.y * nn.y + d_rv.z * nn.z);
      const float t2 = 2.f * f.ndl;
      d_nn = Vec3{t2 * d_rv.x + d_ndl * ld.x, t2 * d_rv.y + d_ndl * ld.y, t2 * d_rv.z + d_ndl * ld.z};
      if (WG) {
        const Vec3 d_ld = {-d_rv.x + d_ndl * nn.x, -d_rv.y + d_ndl * nn.y, -d_rv.z + d_ndl * nn.z};
        const Vec3 dl = normalise_bwd(Vec3{f.lraw[0], f.lraw[1], f.lraw[2]}, d_ld);
        o.g[G_LDE] += dl.x; o.g[G_LDE + 1] += dl.y; o.g[G_LDE + 2] += dl.z;
      }
    }
    if (active) {
      const Vec3 d_normal = normalise_bwd(f.normal, d_nn);
      const float* __restrict__ wen = a.world_to_eye_norm.ptr + (long long)b * a.world_to_eye_norm.batch_stride;
      float wen_acc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // summed over the corners first: 9 updates, not 27
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        d_tc[k] += d_normal.x * f.nvert[k].x + d_normal.y * f.nvert[k].y + d_normal.z * f.nvert[k].z;
        if (WG || WV) {
          const Vec3 d_nv = {tc[k] * d_normal.x, tc[k] * d_normal.y, tc[k] * d_normal.z};
          const Vec3 d_t = normalise_bwd(f.tvert[k], d_nv);
          const Vec3 m = f.mvert[k];
          if (WG) {
            wen_acc[0] += d_t.x * m.x; wen_acc[1] += d_t.x * m.y; wen_acc[2] += d_t.x * m.z;
            wen_acc[3] += d_t.y * m.x; wen_acc[4] += d_t.y * m.y; wen_acc[5] += d_t.y * m.z;
            wen_acc[6] += d_t.z * m.x; wen_acc[7] += d_t.z * m.y; wen_acc[8] += d_t.z * m.z;
          }
          if (WV) {
            const Vec3 d_m = {wen[0] * d_t.x + wen[4] * d_t.y + wen[8] * d_t.z,
                              wen[1] * d_t.x + wen[5] * d_t.y + wen[9] * d_t.z,
                              wen[2] * d_t.x + wen[6] * d_t.y + wen[10] * d_t.z};
            const Vec3 q = normalise3(f.nraw[k]);
            const Vec3 d_q = normalise_bwd(q, d_m);
            const Vec3 dn = normalise_bwd(f.nraw[k], d_q);
            o.d_nrm[k][0] = dn.x; o.d_nrm[k][1] = dn.y; o.d_nrm[k][2] = dn.z;
          }
        }
      }
      if (WG) {
#pragma unroll
        for (int j = 0; j < 9; ++j) o.g[G_WEN + j] += wen_acc[j];
      }
    }
  }

  // ---- common: weights, depth, triangle setup, camera
  if (WG || WV) {
    const float* __restrict__ vp = a.viewport.ptr + (long long)b * a.viewport.batch_stride;
    const float* __restrict__ w2c = a.world_to_clip.ptr + (long long)b * a.world_to_clip.batch_stride;
    const float d_z = d_zw * vp[10];
    if (WG) { o.g[G_VP22] += d_zw * f.z; o.g[G_VP23] += d_zw; }
    const float s = d_tc[0] * tc[0] + d_tc[1] * tc[1] + d_tc[2] * tc[2];
    float d_cc[3], d_zc[3], gv[3];
    // (Reverse-mode arithmetic decides nothing discrete: shared reciprocals instead of IEEE quotients.  The quotients
    // also ran through the compiler's out-of-line division slow path whenever a cotangent was exactly zero -- seven
    // calls per pixel-warp, 14 % of this kernel's instructions.)
    const float r_w = 1.0f / f.w_rec;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      d_cc[j] = (d_tc[j] - s) * r_w + d_z * f.cl[j][2];
      d_zc[j] = d_z * f.cc[j];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
      gv[r] = f.inv[3 * r] * d_cc[0] + f.inv[3 * r + 1] * d_cc[1] + f.inv[3 * r + 2] * d_cc[2];
    if (WG) {
      const float r_vp0 = 1.0f / vp[0], r_vp5 = 1.0f / vp[5];
      o.g[G_VP03] += -gv[0] * r_vp0;
      o.g[G_VP00] += -gv[0] * f.xn * r_vp0;
      o.g[G_VP13] += -gv[1] * r_vp5;
      o.g[G_VP11] += -gv[1] * f.yn * r_vp5;
    }
    if (WG) {
      // d world_to_clip = sum_k (d gl_Position_k) (x) [P_k, 1] with d gl_Position_k = cc_k * (-gv0, -gv1, d_z, -gv2):
      // the outer product factors into g (x) q, q = sum_k cc_k [P_k, 1] -- 16 accumulator updates instead of 48
      const float g4[4] = {-gv[0], -gv[1], d_z, -gv[2]};
      const float q4[4] = {f.cc[0] * f.P[0].x + f.cc[1] * f.P[1].x + f.cc[2] * f.P[2].x,
                           f.cc[0] * f.P[0].y + f.cc[1] * f.P[1].y + f.cc[2] * f.P[2].y,
                           f.cc[0] * f.P[0].z + f.cc[1] * f.P[1].z + f.cc[2] * f.P[2].z,
                           f.cc[0] + f.cc[1] + f.cc[2]};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) o.g[G_W2C + 4 * r + c] += g4[r] * q4[c];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      // d gl_Position of vertex k: (x, y, z, w)
      const float dc[4] = {-f.cc[k] * gv[0], -f.cc[k] * gv[1], d_zc[k], -f.cc[k] * gv[2]};
      if (WV) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          o.d_pos[k][c] = w2c[c] * dc[0] + w2c[4 + c] * dc[1] + w2c[8 + c] * dc[2] + w2c[12 + c] * dc[3];
      }
    }
  }
}

__device__ __forceinline__ void load_cotangent(const JrGradArgs& g, long long gi, bool has_canvas, float& d_zw,
                                               float d_col[3]) {
  d_zw = g.d_zbuffer ? g.d_zbuffer[gi] : 0.f;
  if (has_canvas && g.d_canvas) {
    d_col[0] = g.d_canvas[gi * 3]; d_col[1] = g.d_canvas[gi * 3 + 1]; d_col[2] = g.d_canvas[gi * 3 + 2];
  } else {
    d_col[0] = d_col[1] = d_col[2] = 0.f;
  }
}

// ------------------------------------------------------------ global parameters
constexpr int BWD_THREADS = 256;
#ifndef JR_BWD_MIN_BLOCKS
#define JR_BWD_MIN_BLOCKS 3  // register cap of k_bwd_global (accumulators live in shared memory)
#endif

// Per-pixel outputs of the pixel pass for the one-entry-per-pixel keyed targets (diffuse texture,
// specular map, normal map): key + value are produced HERE, while the pixel's fragment is live, so the
// keyed passes only sort and reduce -- they do not shade the pixel again.
struct PixelEmit {
  unsigned* iota;       // (B*W*H) payload = pixel index (sort values)
  unsigned* key_tex;    // texel key (shared by texture and normal map), or null
  unsigned* key_spec;   // specular-map key, or null
  float* val_tex;       // (B*W*H,3) d colour / d texel
  float* val_spec;      // (B*W*H)   d colour / d specular exponent
  float* val_nmap;      // (B*W*H,3) d colour / d normal-map texel (phong_darboux)
  unsigned inv_tex, inv_spec;            // "no contribution" keys
  long long per_image_tex, per_image_spec;  // key offset per image when the target is batched, else 0
  // three-entries-per-pixel targets (one per triangle corner): values only, keys come from k_bwd_keys
  float* val_pos;       // (B*W*H,3,pos_c) d / d position (+ d / d vertex colour for gouraud: pos_c = 6)
  float* val_nrm;       // (B*W*H,3,3)
  float* val_uv;        // (B*W*H,3,2)  phong_darboux
  float* val_pos2;      // (B*W*H,3,3)  phong_darboux: tangent-frame triangle
  int pos_c;
};

// Extended records of the visible triangles (see rec_store): the vertex stage runs once per triangle
// instead of once per covered pixel of the pixel pass.
template <int S>
__global__ void __launch_bounds__(128) k_bwd_records(const __grid_constant__ JrRenderArgs a, float* __restrict__ recs,
                                                     const int* __restrict__ list, const int* __restrict__ count) {
  __shared__ __align__(16) float stage[128 * TE_FLOATS];
  __shared__ int s_tri[128];
  const int b = blockIdx.y;
  const int n_vis = count[b];
  const int i0 = blockIdx.x * 128;
  if (i0 >= n_vis) return;
  const int i = i0 + threadIdx.x;
  int t = -1;
  if (i < n_vis) {
    t = list[(long long)b * a.T + i];
    Frag f;
    frag_vertex<S>(a, b, t, f);
    rec_store<S>(f, stage + threadIdx.x * TE_FLOATS);
  }
  s_tri[threadIdx.x] = t;
  __syncthreads();
  constexpr int Q = TE_FLOATS / 4;
  const int n = min(128, n_vis - i0) * Q;
  const float4* src = reinterpret_cast<const float4*>(stage);
  float4* base = reinterpret_cast<float4*>(recs + (size_t)b * a.T * TE_FLOATS);
  for (int j = threadIdx.x; j < n; j += 128) {
    const int r = j / Q;
    base[(size_t)s_tri[r] * Q + (j - r * Q)] = src[j];
  }
}

template <int S, bool WV, bool REC>
__global__ void __launch_bounds__(BWD_THREADS, JR_BWD_MIN_BLOCKS)
k_bwd_global(const __grid_constant__ JrRenderArgs a, const __grid_constant__ JrGradArgs g, float* __restrict__ partials,
             const __grid_constant__ PixelEmit em, const float* __restrict__ recs) {
  extern __shared__ float s_acc[];  // [NG][BWD_THREADS]
  const int b = blockIdx.y;
  const int npix = a.W * a.H;
  PixGrad o;
  o.g.p = s_acc + threadIdx.x;
#pragma unroll
  for (int j = 0; j < NG; ++j) o.g[j] = 0.f;
  for (int pix = blockIdx.x * BWD_THREADS + threadIdx.x; pix < npix; pix += gridDim.x * BWD_THREADS) {
    const long long gi = (long long)b * npix + pix;
    const int tri = a.tri_id[gi];
    if (em.iota) em.iota[gi] = (unsigned)gi;
    if (tri < 0) {
      if (em.key_tex) em.key_tex[gi] = em.inv_tex;
      if (em.key_spec) em.key_spec[gi] = em.inv_spec;
      continue;
    }
    const int x = pix / a.H, y = pix - x * a.H;
    Frag f;
    if (REC) {
      rec_load<S>(a, b, recs + ((size_t)b * a.T + tri) * TE_FLOATS, f);
      frag_pixel<S>(a, b, x, y, f);
    } else {
      shade_pixel<S>(a, b, x, y, tri, f);
    }
    float d_zw, d_col[3];
    load_cotangent(g, gi, S != JR_DEPTH, d_zw, d_col);
    backprop_pixel<S, true, true, WV>(a, b, f, d_zw, d_col, o);
    if (WV) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (em.val_pos) {
          float* dst = em.val_pos + (gi * 3 + k) * em.pos_c;
          dst[0] = o.d_pos[k][0]; dst[1] = o.d_pos[k][1]; dst[2] = o.d_pos[k][2];
          if (S == JR_GOURAUD && em.pos_c == 6) { dst[3] = o.d_col[k][0]; dst[4] = o.d_col[k][1]; dst[5] = o.d_col[k][2]; }
        }
        if (S != JR_DEPTH && em.val_nrm) {
          float* dst = em.val_nrm + (gi * 3 + k) * 3;
          dst[0] = o.d_nrm[k][0]; dst[1] = o.d_nrm[k][1]; dst[2] = o.d_nrm[k][2];
        }
        if (S == JR_PHONG_DARBOUX) {
          if (em.val_uv) { em.val_uv[(gi * 3 + k) * 2] = o.d_uv[k][0]; em.val_uv[(gi * 3 + k) * 2 + 1] = o.d_uv[k][1]; }
          if (em.val_pos2) {
            float* dst = em.val_pos2 + (gi * 3 + k) * 3;
            dst[0] = o.d_pos2[k][0]; dst[1] = o.d_pos2[k][1]; dst[2] = o.d_pos2[k][2];
          }
        }
      }
    }
    if (S >= JR_GOURAUD_TEXTURE) {
      if (em.key_tex) {
        bool contributes = f.texel >= 0;
        if (S == JR_PHONG || S == JR_PHONG_DARBOUX) contributes = contributes && f.ok;
        em.key_tex[gi] = contributes ? (unsigned)(b * em.per_image_tex + f.texel) : em.inv_tex;
        if (em.val_tex) { em.val_tex[gi * 3] = o.d_tex[0]; em.val_tex[gi * 3 + 1] = o.d_tex[1]; em.val_tex[gi * 3 + 2] = o.d_tex[2]; }
        if (S == JR_PHONG_DARBOUX && em.val_nmap) {
          em.val_nmap[gi * 3] = o.d_nmap[0]; em.val_nmap[gi * 3 + 1] = o.d_nmap[1]; em.val_nmap[gi * 3 + 2] = o.d_nmap[2];
        }
      }
      if (S >= JR_PHONG_REFLECTION && em.key_spec) {
        em.key_spec[gi] = f.spec_idx >= 0 ? (unsigned)(b * em.per_image_spec + f.spec_idx) : em.inv_spec;
        em.val_spec[gi] = o.d_sexp;
      }
    }
  }
  // fixed-shape block reduction: xor-butterfly inside each warp, then warps in order
  __shared__ float red[BWD_THREADS / 32][NG];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < NG; ++j) {
    float v = o.g[j];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) red[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < NG) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < BWD_THREADS / 32; ++w) v += red[w][threadIdx.x];
    partials[((long long)b * gridDim.x + blockIdx.x) * NG + threadIdx.x] = v;
  }
}

struct GlobalOut {
  float* ptr[NG];
  long long stride[NG];
};

// One CTA per slot.  Shared (stride 0) parameters: the B*nblk partials are summed by 128 threads in a
// fixed strided order, then a fixed-shape tree.  Batched parameters: thread t sums the nblk partials
// of images t, t+128, ... in block order.  Deterministic either way.
__global__ void __launch_bounds__(128) k_bwd_global_final(const float* __restrict__ partials, int B, int nblk,
                                                          GlobalOut out) {
  const int j = blockIdx.x;
  if (out.ptr[j] == nullptr) return;
  const int tid = threadIdx.x;
  if (out.stride[j] == 0) {
    __shared__ float red[128];
    const long long n = (long long)B * nblk;
    float acc = 0.f;
    for (long long i = tid; i < n; i += 128) acc += partials[i * NG + j];
    red[tid] = acc;
    __syncthreads();
    for (int off = 64; off > 0; off >>= 1) {
      if (tid < off) red[tid] += red[tid + off];
      __syncthreads();
    }
    if (tid == 0) *out.ptr[j] += red[0];
  } else {
    for (int b = tid; b < B; b += 128) {
      float acc = 0.f;
      for (int k = 0; k < nblk; ++k) acc += partials[((long long)b * nblk + k) * NG + j];
      out.ptr[j][(long long)b * out.stride[j]] += acc;
    }
  }
}

// --------------------------------------------------------------- keyed targets
struct KeyedPlan {
  int mode;
  int per_pixel;        // entries per pixel (1 or 3)
  long long n_entries;  // B*W*H*per_pixel
  long long keys_per_image;
  bool batched;         // target has a batch axis
  unsigned invalid_key;
  int C;                // channels reduced
  float* out;           // target base
  float* out2;          // MODE_POS: d_colour (may be null)
  const float* pix_vals;      // one-entry-per-pixel modes: values emitted by the pixel pass (C per pixel)
  const unsigned* pix_keys;   // ... and their keys (sort input); payload = PixelEmit::iota
};

template <int S, int MODE>
__global__ void __launch_bounds__(256)
k_bwd_keys(const __grid_constant__ JrRenderArgs a, KeyedPlan plan, unsigned* __restrict__ keys,
           unsigned* __restrict__ vals) {
  const int npix = a.W * a.H;
  for (int b = blockIdx.y; b < a.B; b += gridDim.y)
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
    const long long gi = (long long)b * npix + pix;
    const int tri = a.tri_id[gi];
    const unsigned boff = plan.batched ? (unsigned)((long long)b * plan.keys_per_image) : 0u;
    {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        unsigned key = plan.invalid_key;
        if (tri >= 0) {
          const int32_t* fp = a.faces.ptr + (long long)b * a.faces.batch_stride + 3 * tri;
          if (MODE == MODE_NRM && a.faces_norm.ptr)
            fp = a.faces_norm.ptr + (long long)b * a.faces_norm.batch_stride + 3 * tri;
          int v = fp[k];
          if (MODE == MODE_UV || MODE == MODE_POS2) {
            // vertices of the tangent-frame triangle of the chosen triangle's first vertex
            const int v0 = min(max(fp[0], 0), a.n_pos - 1);
            // wrap one negative, then clamp: the forward's gathers (jr_shade.cuh, frag_pixel Darboux branch)
            const int face = wrap_clamp((a.id_to_face.ptr + (long long)b * a.id_to_face.batch_stride)[v0], a.n_faces_indices);
            v = wrap_clamp((a.faces_indices.ptr + (long long)b * a.faces_indices.batch_stride)[3 * face + k],
                           (int)plan.keys_per_image);
          }
          v = min(max(v, 0), (int)plan.keys_per_image - 1);
          key = boff + (unsigned)v;
        }
        keys[gi * 3 + k] = key;
        vals[gi * 3 + k] = (unsigned)(gi * 4 + k);
      }
    }
  }
}

template <int C>
struct Carry {
  unsigned first_key, last_key;
  int first_open, last_open, whole, pad;
  float first_val[C], last_val[C];
};

// One thread per sorted entry (it gathers the value the pixel pass emitted for its pixel / corner),
// fixed chunks of 256 entries per CTA.
template <int S, int MODE, int C>
__global__ void __launch_bounds__(256)
k_bwd_segreduce(const __grid_constant__ JrRenderArgs a, const __grid_constant__ JrGradArgs g, KeyedPlan plan,
                const unsigned* __restrict__ keys, const unsigned* __restrict__ vals, Carry<C>* __restrict__ carry) {
  __shared__ unsigned s_key[256];
  __shared__ int s_head[256];
  __shared__ float s_val[C][256];
  const int tid = threadIdx.x;
  const long long i = (long long)blockIdx.x * 256 + tid;
  unsigned key = plan.invalid_key;
  float v[C];
#pragma unroll
  for (int c = 0; c < C; ++c) v[c] = 0.f;
  if (i < plan.n_entries) {
    key = keys[i];
    if (key != plan.invalid_key && plan.pix_vals) {
      const unsigned payload = vals[i];
      // one entry per pixel: payload = pixel; one per corner: payload = pixel * 4 + corner
      const long long e = (plan.per_pixel == 1) ? (long long)payload : (long long)(payload >> 2) * 3 + (payload & 3);
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] = plan.pix_vals[e * C + c];
    }
  }
  s_key[tid] = key;
#pragma unroll
  for (int c = 0; c < C; ++c) s_val[c][tid] = v[c];
  __syncthreads();
  // head index of my segment: max-scan of head positions
  const bool head = (tid == 0) || (s_key[tid - 1] != key);
  s_head[tid] = head ? tid : 0;
  __syncthreads();
  for (int off = 1; off < 256; off <<= 1) {
    int t = (tid >= off) ? s_head[tid - off] : 0;
    __syncthreads();
    s_head[tid] = max(s_head[tid], t);
    __syncthreads();
  }
  const int my_head = s_head[tid];
  // segmented inclusive scan (Hillis-Steele: fixed association -> deterministic)
  for (int off = 1; off < 256; off <<= 1) {
    float t[C];
    const bool take = (tid >= off) && (tid - off >= my_head);
#pragma unroll
    for (int c = 0; c < C; ++c) t[c] = take ? s_val[c][tid - off] : 0.f;
    __syncthreads();
    if (take) {
#pragma unroll
      for (int c = 0; c < C; ++c) s_val[c][tid] += t[c];
    }
    __syncthreads();
  }
  const bool tail = (tid == 255) || (s_key[tid + 1] != key);
  if (!tail) return;
  // neighbours across the chunk boundary
  const long long c0 = (long long)blockIdx.x * 256;
  const bool open_left = (my_head == 0) && (c0 > 0) && (keys[c0 - 1] == key);
  const bool open_right = (tid == 255) && (c0 + 256 < plan.n_entries) && (keys[c0 + 256] == key);
  Carry<C>& cr = carry[blockIdx.x];
  if (my_head == 0) {
    cr.first_key = key; cr.first_open = open_left ? 1 : 0;
    cr.whole = (tid == 255) ? 1 : 0;
#pragma unroll
    for (int c = 0; c < C; ++c) cr.first_val[c] = s_val[c][tid];
  }
  if (tid == 255) {
    cr.last_key = key; cr.last_open = open_right ? 1 : 0;
#pragma unroll
    for (int c = 0; c < C; ++c) cr.last_val[c] = s_val[c][tid];
  }
  if (key == plan.invalid_key) return;
  if (!open_left && !open_right) {
    // complete segment: single writer of this key in the whole launch
    if (MODE == MODE_POS && C > 3) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        plan.out[(long long)key * 3 + c] += s_val[c][tid];
        if (plan.out2) plan.out2[(long long)key * 3 + c] += s_val[3 + c][tid];
      }
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) plan.out[(long long)key * C + c] += s_val[c][tid];
    }
  }
}

// Resolve segments that straddle chunks: one WARP per chunk; the warp whose chunk STARTS a run finds
// the run's last chunk by binary search in the sorted keys, its lanes sum the per-chunk pieces in a
// fixed strided order and a fixed butterfly combines them (deterministic, and parallel for the long
// runs produced by few-key targets such as 1x1 textures).
template <int MODE, int C>
__global__ void __launch_bounds__(256)
k_bwd_segfix(KeyedPlan plan, const Carry<C>* __restrict__ carry, const unsigned* __restrict__ keys, int nchunks) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= nchunks) return;
  const Carry<C>& me = carry[c];
  if (!me.last_open) return;
  if (me.whole && me.first_open) return;  // continues a run started further left
  const unsigned key = me.last_key;
  if (key == plan.invalid_key) return;
  // last entry with this key: upper bound over the sorted keys
  long long lo = (long long)(c + 1) * 256, hi = plan.n_entries;  // keys[lo] == key is known
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (keys[mid] <= key) lo = mid + 1; else hi = mid;
  }
  const int c_end = (int)((lo - 1) >> 8);  // chunk holding the last entry of the run (> c)
  float acc[C];
#pragma unroll
  for (int k = 0; k < C; ++k) acc[k] = 0.f;
  // (unrolled: the loads of 8 iterations are in flight together -- a shared 1x1 texture gives runs of
  // tens of thousands of chunks, and one round trip per iteration made this kernel latency-bound)
#pragma unroll 8
  for (int j = c + 1 + lane; j <= c_end; j += 32) {
    const Carry<C>& nx = carry[j];
#pragma unroll
    for (int k = 0; k < C; ++k) acc[k] += nx.first_val[k];
  }
#pragma unroll
  for (int k = 0; k < C; ++k) {
    float v = acc[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    acc[k] = v + me.last_val[k];
  }
  if (lane != 0) return;
  if (MODE == MODE_POS && C > 3) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      plan.out[(long long)key * 3 + k] += acc[k];
      if (plan.out2) plan.out2[(long long)key * 3 + k] += acc[3 + k];
    }
  } else {
#pragma unroll
    for (int k = 0; k < C; ++k) plan.out[(long long)key * C + k] += acc[k];
  }
}

// d(out)/d(old buffer) = 1 - keep (pipeline.py:420-437)
__global__ void k_bwd_mask(const int32_t* __restrict__ tri_id, float* d_z, float* d_c, long long total) {
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < total;
       gi += (long long)gridDim.x * blockDim.x) {
    if (tri_id[gi] >= 0) {
      if (d_z) d_z[gi] = 0.f;
      if (d_c) { d_c[gi * 3] = 0.f; d_c[gi * 3 + 1] = 0.f; d_c[gi * 3 + 2] = 0.f; }
    }
  }
}

// ----------------------------------------------------------------- host side
static int bit_length(unsigned long long v) { int n = 0; while (v) { ++n; v >>= 1; } return n; }
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static const bool g_no_bwd_rec = getenv("JR_NO_BWD_REC") != nullptr;  // A/B switch: pixel pass recomputes the vertex stage

struct BwdLayout {
  int nblk;             // global pass blocks per image
  size_t partials, keys_a, keys_b, vals_a, vals_b, cub_temp, carry, total;
  size_t em_iota, em_key_tex, em_key_spec, em_val_tex, em_val_spec, em_val_nmap;  // PixelEmit buffers
  size_t em_val_pos, em_val_nrm, em_val_uv, em_val_pos2;
  int pos_c;
  bool use_rec;
  size_t rec_off, rec_flags, rec_list;  // extended records of the visible triangles + their list
  size_t cub_bytes;
  long long max_entries;
};

static bool wants_global(const JrGradArgs* g) {
  return g->d_world_to_clip.ptr || g->d_viewport.ptr || g->d_world_to_eye_norm.ptr || g->d_light_direction.ptr ||
         g->d_light_colour.ptr || g->d_light_dir_eye.ptr || g->d_ambient.ptr || g->d_diffuse.ptr ||
         g->d_specular.ptr || g->d_shadow_strength.ptr;
}

static BwdLayout bwd_layout(const JrRenderArgs* a, const JrGradArgs* g) {
  BwdLayout L{};
  const long long npix = (long long)a->W * a->H;
  // blocks per image of the pixel pass: 8 to 32 pixels per thread -- few enough blocks that the 52-slot block
  // reduction at the end is amortised (it cost ~5 % of the pass at 7 pixels per thread), enough of them
  // (about 2048 CTAs in total) to fill the GPU when the batch is small
  const int nblk_max = (int)((npix + BWD_THREADS * 8 - 1) / (BWD_THREADS * 8));
  const int nblk_min = (int)((npix + BWD_THREADS * 32 - 1) / (BWD_THREADS * 32));
  L.nblk = (2048 + a->B - 1) / a->B;
  if (L.nblk > nblk_max) L.nblk = nblk_max;
  if (L.nblk < nblk_min) L.nblk = nblk_min;
  if (L.nblk < 1) L.nblk = 1;
  if (L.nblk > 64) L.nblk = 64;
  const bool keyed1 = g->d_texture.ptr || g->d_specular_map.ptr || g->d_normal_map.ptr;
  const bool keyed3 = g->d_position.ptr || g->d_colour.ptr || g->d_normal.ptr || g->d_uv.ptr;
  L.max_entries = keyed3 ? npix * a->B * 3 : (keyed1 ? npix * a->B : 0);
  size_t off = 0;
  L.partials = off; off += align256(sizeof(float) * NG * (size_t)L.nblk * a->B);
  if (L.max_entries > 0) {
    const size_t n = (size_t)L.max_entries;
    L.keys_a = off; off += align256(n * 4);
    L.keys_b = off; off += align256(n * 4);
    L.vals_a = off; off += align256(n * 4);
    L.vals_b = off; off += align256(n * 4);
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const unsigned*)nullptr, (unsigned*)nullptr,
                                    (const unsigned*)nullptr, (unsigned*)nullptr, (int)n, 0, 32);
    L.cub_bytes = tmp;
    L.cub_temp = off; off += align256(tmp);
    const size_t nchunks = (n + 255) / 256;
    L.carry = off; off += align256(nchunks * sizeof(Carry<6>));
  }
  if (keyed1) {
    const size_t np = (size_t)npix * a->B;
    L.em_iota = off; off += align256(np * 4);
    if (g->d_texture.ptr || g->d_normal_map.ptr) { L.em_key_tex = off; off += align256(np * 4); }
    if (g->d_texture.ptr) { L.em_val_tex = off; off += align256(np * 12); }
    if (g->d_normal_map.ptr) { L.em_val_nmap = off; off += align256(np * 12); }
    if (g->d_specular_map.ptr) {
      L.em_key_spec = off; off += align256(np * 4);
      L.em_val_spec = off; off += align256(np * 4);
    }
  }
  if (keyed3) {
    const size_t np = (size_t)npix * a->B;
    L.pos_c = (a->shader == JR_GOURAUD && g->d_colour.ptr) ? 6 : 3;
    if (g->d_position.ptr || g->d_colour.ptr) { L.em_val_pos = off; off += align256(np * 3 * L.pos_c * 4); }
    if (g->d_normal.ptr) { L.em_val_nrm = off; off += align256(np * 36); }
    if (a->shader == JR_PHONG_DARBOUX) {
      if (g->d_uv.ptr) { L.em_val_uv = off; off += align256(np * 24); }
      if (g->d_position.ptr) { L.em_val_pos2 = off; off += align256(np * 36); }
    }
  }
  // records pay off when a triangle is shared by several pixels (the forward's rule for attribute records)
  L.use_rec = (a->shader == JR_PHONG_REFLECTION || a->shader == JR_PHONG_REFLECTION_SHADOW) && a->T > 0 &&
              npix >= 2LL * a->T && !g_no_bwd_rec;
  if (L.use_rec) {
    L.rec_off = off; off += align256((size_t)a->B * a->T * TE_FLOATS * 4);
    L.rec_flags = off; off += align256((((size_t)a->B * a->T + 31) / 32) * 4 + (size_t)a->B * 4);
    L.rec_list = off; off += align256((size_t)a->B * a->T * 4);
  }
  L.total = off;
  return L;
}

template <int S, int MODE, int C>
static int run_keyed(const JrRenderArgs* a, const JrGradArgs* g, const BwdLayout& L, KeyedPlan plan,
                     cudaStream_t stream) {
  char* ws = (char*)g->workspace;
  unsigned* keys_a = (unsigned*)(ws + L.keys_a);
  unsigned* keys_b = (unsigned*)(ws + L.keys_b);
  unsigned* vals_a = (unsigned*)(ws + L.vals_a);
  unsigned* vals_b = (unsigned*)(ws + L.vals_b);
  Carry<C>* carry = (Carry<C>*)(ws + L.carry);
  int bx = (a->W * a->H + 255) / 256;
  if (bx > 4096) bx = 4096;
  const unsigned* keys_in = keys_a;
  const unsigned* vals_in = vals_a;
  if (plan.pix_keys) {  // emitted by the pixel pass
    keys_in = plan.pix_keys;
    vals_in = (const unsigned*)(ws + L.em_iota);
  } else {
    k_bwd_keys<S, MODE><<<dim3(bx, a->B > 65535 ? 65535 : a->B), 256, 0, stream>>>(*a, plan, keys_a, vals_a);
    g_launches++;
    mark(stream, "k_bwd_keys");
  }
  size_t tmp = L.cub_bytes;
  const int end_bit = bit_length((unsigned long long)plan.invalid_key);
  cub::DeviceRadixSort::SortPairs(ws + L.cub_temp, tmp, keys_in, keys_b, vals_in, vals_b, (int)plan.n_entries, 0,
                                  end_bit, stream);
  mark(stream, "cub::DeviceRadixSort");
  const long long nchunks = (plan.n_entries + 255) / 256;
  k_bwd_segreduce<S, MODE, C><<<(unsigned)nchunks, 256, 0, stream>>>(*a, *g, plan, keys_b, vals_b, carry);
  g_launches++;
  mark(stream, "k_bwd_segreduce");
  k_bwd_segfix<MODE, C><<<(unsigned)((nchunks + 7) / 8), 256, 0, stream>>>(plan, carry, keys_b, (int)nchunks);
  g_launches++;
  mark(stream, "k_bwd_segfix");
  return cudaGetLastError() == cudaSuccess ? JR_OK : JR_ERR_CUDA;
}

template <int S>
static int backward_impl(const JrRenderArgs* a, const JrGradArgs* g, cudaStream_t stream) {
  const BwdLayout L = bwd_layout(a, g);
  if (L.total > 0 && (!g->workspace || g->workspace_bytes < L.total)) return JR_ERR_WORKSPACE;
  char* ws = (char*)g->workspace;
  const long long npix = (long long)a->W * a->H;
  const long long total = npix * a->B;
  if (total * 4 > 0xFFFFFFFFLL) return JR_ERR_DIMS;  // payload packing (gi * 4 + corner) is 32-bit
  int rc = JR_OK;
  mark(stream, nullptr);
  // ---- pixel pass: scene-global partial sums + keys / values of the one-entry-per-pixel targets
  const bool want_tex = S >= JR_GOURAUD_TEXTURE && g->d_texture.ptr;
  const bool want_spec = S >= JR_PHONG_REFLECTION && g->d_specular_map.ptr;
  const bool want_nmap = S == JR_PHONG_DARBOUX && g->d_normal_map.ptr;
  PixelEmit em{};
  long long nk_tex = 0, nk_spec = 0;
  if (want_tex || want_nmap) {
    const long long per = (long long)a->tex_w * a->tex_h;
    const long long bs_t = g->d_texture.ptr ? g->d_texture.batch_stride : 0;
    const long long bs_n = g->d_normal_map.ptr ? g->d_normal_map.batch_stride : 0;
    if (want_tex && want_nmap && ((bs_t != 0) != (bs_n != 0))) return JR_ERR_UNSUPPORTED;  // one key space
    const bool batched = (want_tex ? bs_t : bs_n) != 0;
    nk_tex = per * (batched ? a->B : 1);
    if (nk_tex >= 0x7FFFFFFFLL) return JR_ERR_DIMS;
    em.key_tex = (unsigned*)(ws + L.em_key_tex);
    em.inv_tex = (unsigned)nk_tex;
    em.per_image_tex = batched ? per : 0;
    if (want_tex) em.val_tex = (float*)(ws + L.em_val_tex);
    if (want_nmap) em.val_nmap = (float*)(ws + L.em_val_nmap);
  }
  if (want_spec) {
    const long long per = (long long)a->spec_w * a->spec_h;
    const bool batched = g->d_specular_map.batch_stride != 0;
    nk_spec = per * (batched ? a->B : 1);
    if (nk_spec >= 0x7FFFFFFFLL) return JR_ERR_DIMS;
    em.key_spec = (unsigned*)(ws + L.em_key_spec);
    em.val_spec = (float*)(ws + L.em_val_spec);
    em.inv_spec = (unsigned)nk_spec;
    em.per_image_spec = batched ? per : 0;
  }
  if (want_tex || want_nmap || want_spec) em.iota = (unsigned*)(ws + L.em_iota);
  const bool want_pos = g->d_position.ptr || (S == JR_GOURAUD && g->d_colour.ptr);
  const bool want_nrm = S != JR_DEPTH && g->d_normal.ptr;
  const bool want_uv = S == JR_PHONG_DARBOUX && g->d_uv.ptr;
  const bool want_vertex = want_pos || want_nrm || want_uv;
  if (want_pos) {
    if (!g->d_position.ptr) return JR_ERR_UNSUPPORTED;  // d_colour alone: ask for d_position too
    em.val_pos = (float*)(ws + L.em_val_pos);
    em.pos_c = L.pos_c;
    if (S == JR_PHONG_DARBOUX) em.val_pos2 = (float*)(ws + L.em_val_pos2);
  }
  if (want_nrm) em.val_nrm = (float*)(ws + L.em_val_nrm);
  if (want_uv) em.val_uv = (float*)(ws + L.em_val_uv);
  if (wants_global(g) || em.iota || want_vertex) {
    float* partials = (float*)(ws + L.partials);
    dim3 grid(L.nblk, a->B);
    const float* recs = nullptr;
    constexpr bool CAN_REC = (S == JR_PHONG_REFLECTION || S == JR_PHONG_REFLECTION_SHADOW);
    if (CAN_REC && L.use_rec) {
      unsigned* flag_words = (unsigned*)(ws + L.rec_flags);
      const size_t n_words = ((size_t)a->B * a->T + 31) / 32;
      int* count = (int*)(flag_words + n_words);
      int* list = (int*)(ws + L.rec_list);
      cudaMemsetAsync(flag_words, 0, n_words * 4 + (size_t)a->B * 4, stream);
      int bx = (int)((npix + 255) / 256);
      if (bx > 64) bx = 64;
      k_mark_visible<0><<<dim3(bx, a->B), 256, 0, stream>>>(a->tri_id, flag_words, list, count, (int)npix, a->T, a->B);
      k_bwd_records<CAN_REC ? S : JR_PHONG_REFLECTION><<<dim3((a->T + 127) / 128, a->B), 128, 0, stream>>>(
          *a, (float*)(ws + L.rec_off), list, count);
      g_launches += 2;
      mark(stream, "k_mark_vis+k_bwd_recs");
      recs = (const float*)(ws + L.rec_off);
    }
    const size_t sm = NG * BWD_THREADS * 4;
#define JR_PIXEL_PASS(WV_, REC_)                                                                                 \
  do {                                                                                                            \
    cudaFuncSetAttribute(k_bwd_global<S, WV_, REC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);       \
    k_bwd_global<S, WV_, REC_><<<grid, BWD_THREADS, sm, stream>>>(*a, *g, partials, em, recs);                    \
  } while (0)
    if (CAN_REC && recs) {
      if (want_vertex) JR_PIXEL_PASS(true, CAN_REC); else JR_PIXEL_PASS(false, CAN_REC);
    } else {
      if (want_vertex) JR_PIXEL_PASS(true, false); else JR_PIXEL_PASS(false, false);
    }
#undef JR_PIXEL_PASS
    g_launches++;
    mark(stream, "k_bwd_global");
  }
  if (wants_global(g)) {
    float* partials = (float*)(ws + L.partials);
    GlobalOut out{};
    auto set = [&](int base, int n, const JrF32Out& t, const int* offs) {
      for (int j = 0; j < n; ++j) {
        out.ptr[base + j] = t.ptr ? t.ptr + (offs ? offs[j] : j) : nullptr;
        out.stride[base + j] = t.batch_stride;
      }
    };
    set(G_W2C, 16, g->d_world_to_clip, nullptr);
    const int vp_offs[6] = {0, 3, 5, 7, 10, 11};
    set(G_VP00, 6, g->d_viewport, vp_offs);
    const int wen_offs[9] = {0, 1, 2, 4, 5, 6, 8, 9, 10};
    set(G_WEN, 9, g->d_world_to_eye_norm, wen_offs);
    set(G_LDIR, 3, g->d_light_direction, nullptr);
    set(G_LCOL, 3, g->d_light_colour, nullptr);
    set(G_LDE, 3, g->d_light_dir_eye, nullptr);
    set(G_AMB, 3, g->d_ambient, nullptr);
    set(G_DIF, 3, g->d_diffuse, nullptr);
    set(G_SPE, 3, g->d_specular, nullptr);
    set(G_STR, 3, g->d_shadow_strength, nullptr);
    k_bwd_global_final<<<NG, 128, 0, stream>>>(partials, a->B, L.nblk, out);
    g_launches++;
    mark(stream, "k_bwd_global_final");
  }
  if (want_tex) {
    KeyedPlan p{};
    p.mode = MODE_TEXEL; p.per_pixel = 1; p.n_entries = total;
    p.keys_per_image = (long long)a->tex_w * a->tex_h;
    p.batched = g->d_texture.batch_stride != 0;
    p.invalid_key = em.inv_tex; p.C = 3; p.out = g->d_texture.ptr; p.out2 = nullptr;
    p.pix_keys = em.key_tex; p.pix_vals = em.val_tex;
    rc = run_keyed<S, MODE_TEXEL, 3>(a, g, L, p, stream);
    if (rc != JR_OK) return rc;
  }
  if (want_spec) {
    KeyedPlan p{};
    p.mode = MODE_SPEC; p.per_pixel = 1; p.n_entries = total;
    p.keys_per_image = (long long)a->spec_w * a->spec_h;
    p.batched = g->d_specular_map.batch_stride != 0;
    p.invalid_key = em.inv_spec; p.C = 1; p.out = g->d_specular_map.ptr; p.out2 = nullptr;
    p.pix_keys = em.key_spec; p.pix_vals = em.val_spec;
    rc = run_keyed<S, MODE_SPEC, 1>(a, g, L, p, stream);
    if (rc != JR_OK) return rc;
  }
  if (g->d_position.ptr || (S == JR_GOURAUD && g->d_colour.ptr)) {
    KeyedPlan p{};
    p.mode = MODE_POS; p.per_pixel = 3; p.n_entries = total * 3;
    p.keys_per_image = a->n_pos;
    const long long bs = g->d_position.ptr ? g->d_position.batch_stride : g->d_colour.batch_stride;
    p.batched = bs != 0;
    const long long nk = p.keys_per_image * (p.batched ? a->B : 1);
    if (nk >= 0x7FFFFFFFLL || p.n_entries >= 0x7FFFFFFFLL) return JR_ERR_DIMS;
    p.invalid_key = (unsigned)nk; p.C = L.pos_c;
    p.out = g->d_position.ptr; p.out2 = (S == JR_GOURAUD) ? g->d_colour.ptr : nullptr;
    p.pix_vals = em.val_pos;
    rc = (L.pos_c == 6) ? run_keyed<S, MODE_POS, 6>(a, g, L, p, stream) : run_keyed<S, MODE_POS, 3>(a, g, L, p, stream);
    if (rc != JR_OK) return rc;
  }
  if (S != JR_DEPTH && g->d_normal.ptr) {
    KeyedPlan p{};
    p.mode = MODE_NRM; p.per_pixel = 3; p.n_entries = total * 3;
    p.keys_per_image = a->n_nrm;
    p.batched = g->d_normal.batch_stride != 0;
    const long long nk = p.keys_per_image * (p.batched ? a->B : 1);
    if (nk >= 0x7FFFFFFFLL || p.n_entries >= 0x7FFFFFFFLL) return JR_ERR_DIMS;
    p.invalid_key = (unsigned)nk; p.C = 3; p.out = g->d_normal.ptr; p.out2 = nullptr;
    p.pix_vals = em.val_nrm;
    rc = run_keyed<S, MODE_NRM, 3>(a, g, L, p, stream);
    if (rc != JR_OK) return rc;
  }
  if (S == JR_PHONG_DARBOUX) {
    if (want_nmap) {
      KeyedPlan p{};
      p.mode = MODE_NMAP; p.per_pixel = 1; p.n_entries = total;
      p.keys_per_image = (long long)a->tex_w * a->tex_h;
      p.batched = g->d_normal_map.batch_stride != 0;
      p.invalid_key = em.inv_tex; p.C = 3; p.out = g->d_normal_map.ptr; p.out2 = nullptr;
      p.pix_keys = em.key_tex; p.pix_vals = em.val_nmap;
      rc = run_keyed<S, MODE_NMAP, 3>(a, g, L, p, stream);
      if (rc != JR_OK) return rc;
    }
    if (g->d_uv.ptr) {
      KeyedPlan p{};
      p.mode = MODE_UV; p.per_pixel = 3; p.n_entries = total * 3;
      p.keys_per_image = a->n_uv;
      p.batched = g->d_uv.batch_stride != 0;
      const long long nk = p.keys_per_image * (p.batched ? a->B : 1);
      if (nk >= 0x7FFFFFFFLL || p.n_entries >= 0x7FFFFFFFLL) return JR_ERR_DIMS;
      p.invalid_key = (unsigned)nk; p.C = 2; p.out = g->d_uv.ptr; p.out2 = nullptr;
      p.pix_vals = em.val_uv;
      rc = run_keyed<S, MODE_UV, 2>(a, g, L, p, stream);
      if (rc != JR_OK) return rc;
    }
    if (g->d_position.ptr) {  // second position pass: the tangent-frame triangle's vertices
      KeyedPlan p{};
      p.mode = MODE_POS2; p.per_pixel = 3; p.n_entries = total * 3;
      p.keys_per_image = a->n_pos;
      p.batched = g->d_position.batch_stride != 0;
      const long long nk = p.keys_per_image * (p.batched ? a->B : 1);
      if (nk >= 0x7FFFFFFFLL || p.n_entries >= 0x7FFFFFFFLL) return JR_ERR_DIMS;
      p.invalid_key = (unsigned)nk; p.C = 3; p.out = g->d_position.ptr; p.out2 = nullptr;
      p.pix_vals = em.val_pos2;
      rc = run_keyed<S, MODE_POS2, 3>(a, g, L, p, stream);
      if (rc != JR_OK) return rc;
    }
  }
  // cotangent of the incoming buffers
  if (!g->no_buffer_grads && (g->d_zbuffer || g->d_canvas)) {
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    k_bwd_mask<<<(unsigned)blocks, 256, 0, stream>>>(a->tri_id, g->d_zbuffer, S != JR_DEPTH ? g->d_canvas : nullptr,
                                                     total);
    g_launches++;
    mark(stream, "k_bwd_mask");
  }
  return cudaGetLastError() == cudaSuccess ? JR_OK : JR_ERR_CUDA;
}

}  // namespace jr

using namespace jr;

extern "C" {

size_t jr_backward_workspace_bytes(const JrRenderArgs* a, const JrGradArgs* g) {
  if (!a || !g || a->B <= 0 || a->W <= 0 || a->H <= 0) return 0;
  return bwd_layout(a, g).total;
}

int jr_render_backward(const JrRenderArgs* a, const JrGradArgs* g, jr_stream_t stream_) {
  if (!a || !g) return JR_ERR_NULL;
  if (a->shader < 0 || a->shader >= JR_NUM_SHADERS) return JR_ERR_SHADER;
  if (a->B <= 0 || a->W <= 0 || a->H <= 0 || a->B > 65535) return JR_ERR_DIMS;  // grid.y = batch
  if (!a->tri_id || !a->world_to_clip.ptr || !a->viewport.ptr) return JR_ERR_NULL;
  if (a->inst_transform.ptr) return JR_ERR_UNSUPPORTED;  // reverse mode wants the materialised world-space arrays
  cudaStream_t stream = (cudaStream_t)stream_;
  switch (a->shader) {
    case JR_DEPTH: return backward_impl<JR_DEPTH>(a, g, stream);
    case JR_GOURAUD: return backward_impl<JR_GOURAUD>(a, g, stream);
    case JR_GOURAUD_TEXTURE: return backward_impl<JR_GOURAUD_TEXTURE>(a, g, stream);
    case JR_PHONG: return backward_impl<JR_PHONG>(a, g, stream);
    case JR_PHONG_REFLECTION: return backward_impl<JR_PHONG_REFLECTION>(a, g, stream);
    case JR_PHONG_REFLECTION_SHADOW: return backward_impl<JR_PHONG_REFLECTION_SHADOW>(a, g, stream);
    case JR_PHONG_DARBOUX:
      if (!a->id_to_face.ptr || !a->faces_indices.ptr || !a->normal_map.ptr || !a->uv.ptr) return JR_ERR_NULL;
      return backward_impl<JR_PHONG_DARBOUX>(a, g, stream);
    default: return JR_ERR_SHADER;
  }
}

}  // extern "C"
