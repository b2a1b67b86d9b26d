// jr_shade.cuh -- per-pixel interpolate + fragment + mix of the seven built-in
// shaders (reference renderer/shaders/*.py), shared by the forward shading kernel
// and the backward kernels so that both see bit-identical discrete decisions
// (texel choice, shadow test, keep).  Operation order mirrors oracle/jr_oracle.py.
#pragma once
#include "../../include/jr_b200.h"
#include "jr_device.cuh"
#include "jr_geometry.cuh"

namespace jr {

__device__ __forceinline__ void load3(const JrF32& arr, int b, float out[3]) {
  const float* p = arr.ptr + (long long)b * arr.batch_stride;
  out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
}
__device__ __forceinline__ Vec3 loadv3(const float* p, int i) {
  return Vec3{p[3 * i], p[3 * i + 1], p[3 * i + 2]};
}

// Everything the fragment stage computed for one pixel (kept for backward).
struct Frag {
  int fi[3], fn[3], fu[3];     // vertex ids of the chosen triangle: position / normal / uv rows
  Vec3 P[3];                   // world positions
  float cl[3][4];              // clip positions (gl_Position)
  float inv[9];                // PerPrimitive.matrix_inv
  float xn, yn;                // pixel in NDC
  float cc[3], w_rec, z, zw;   // clip_coef, 1/w, z_ndc, gl_FragCoord.z
  float tc[3];                 // true_clip_coef (perspective-correct weights)
  float col[3];
  bool keep;
  // shader intermediates
  float lcol[3];
  float inten[3];              // S2,S3: n_k . nl
  Vec3 nraw[3];                // raw per-vertex normals
  Vec3 nvert[3];               // S2,S3: normalise(n_k);  S4-S7: eye-space vertex normals
  Vec3 mvert[3], tvert[3];     // S4-S7: m = normalise(normalise(n_k)), t = R m (before the last normalise)
  Vec3 nl;                     // normalised light direction (S2-S5) / light_dir_eye (S6,S7)
  float lraw[3];               // raw light vector that was normalised into nl
  float colv[3][3];            // S2: per-vertex colours (colour * light colour * intensity)
  float uvv[3][2];             // S3-S7: per-vertex uv
  float scv[3][4];             // S7: per-vertex shadow coordinates (NDC of the light camera)
  int ti;                      // S6,S7: texture index of the triangle (first vertex)
  float lc[3];                 // S3: interpolated light colour; S4,S5: lcol * ndl
  long long texel;             // linear texel index (u * tex_h + v), -1 if none
  long long spec_idx;          // linear index into specular_map (S6,S7)
  float tex[3];
  Vec3 normal, nn;             // interpolated normal and its normalisation
  float ndl;
  bool ok;                     // S4,S5: all(lc >= 0)
  float diffuse, specular, sexp, base;
  Vec3 rv;                     // 2 ndl nn - ld (before normalise)
  float ds[3], shadow[3], amb[3], dif[3], spe[3], str[3];
  bool lit;
  // S5 (Darboux) intermediates, kept for backward
  int dvtx[3];                 // vertices of the tangent-frame triangle (faces_indices[id_to_face[v0]])
  int dvuv[3];                 // the same ids clamped to the uv rows
  Vec3 dP[3];                  // their world positions
  float dtr[3][3], dtw[3];     // their NDC positions and clip w
  float dAI[9];                // inverse of [tr1 - tr0; tr2 - tr0; n]
  float dduv[4];               // du0, du1, dv0, dv1
  Vec3 divr, djvr, div_, djv;  // tangent / bitangent before and after normalisation
  float dnm[3];                // normal-map texel
  Vec3 dbn, dn2;               // B @ nm and its normalisation (the shading normal)
};

// Per-TRIANGLE part of the common setup (independent of the pixel).
__device__ __forceinline__ void frag_setup_tri(const JrRenderArgs& a, int b, int tri, Frag& f) {
  const float* __restrict__ w2c = a.world_to_clip.ptr + (long long)b * a.world_to_clip.batch_stride;
  const float* __restrict__ pos = a.position.ptr + (long long)b * a.position.batch_stride;
  const int32_t* __restrict__ faces = a.faces.ptr + (long long)b * a.faces.batch_stride;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    f.fi[k] = min(max(faces[3 * tri + k], 0), a.n_pos - 1);  // clamped like the visibility kernels
    f.P[k] = fetch_position(a, b, pos, f.fi[k]);
    to_clip(w2c, f.P[k].x, f.P[k].y, f.P[k].z, f.cl[k]);
  }
  float M[9];
  tri_matrix(f.cl[0], f.cl[1], f.cl[2], M);
  lu_inverse3(M, f.inv);
}

// base^e for the Phong lobe: base is a clamped cosine in [0, 1].  2^(e * log2(base)) with the library's log2f / exp2f
// (1 and 2 ulp) is accurate to |L| * 2^L * 1.2e-7 <= 6e-8 ABSOLUTE (L = e * log2(base) <= 0: the relative error grows
// with |L| exactly as fast as the result shrinks) -- the accuracy class of powf itself (which is not correctly rounded
// either: the oracle's pow and CUDA's differ in the last place) -- at ~40 instead of ~95 instructions.  Nothing
// discrete depends on it.  e == 0 gives 1 for every base, base 0 gives 0 / inf for positive / negative e, as pow.
__device__ __forceinline__ float pow_unit(float base, float e) {
  return e == 0.f ? 1.f : exp2f(e * log2f(base));
}

// Per-IMAGE constants of the pixel stage.  The record-based forward path writes them ONCE per image into the
// workspace (k_tri_attr's first block, pix_const_write) -- light / material parameters, the NORMALISED light
// direction (the reference normalises it per fragment: the same operations on the same inputs, done once), the
// viewport entries, the shadow viewport, the reciprocal of H -- and the pixel kernels read them through one
// pointer: a pixel then needs no 64-bit address arithmetic per parameter and skips a sqrt and three IEEE divisions.
// Values are bit-identical to what frag_pixel computes on its own (PC == false).  (Staging them in shared memory
// per CTA was measured slower: two barriers and a dependent load -> sqrt -> divide chain in front of 256 pixels.)
struct __align__(16) PixConst {
  float vp0, vp3, vp5, vp7, vp10, vp11;
  float lcol[3], nl[3], amb[3], dif[3], spe[3], str[3];
  float svp[16];
  unsigned h_magic;   // ceil(2^32 / H) when floor(pix / H) == umulhi(pix, h_magic) for every pixel index, else 0
  unsigned pad0;
  const float* xs;    // the image's pixel -> NDC tables (W and H entries): frag_setup_pix's two divisions, done once
  const float* ys;    // per column / row instead of once per pixel
  float pad[2];
};
static_assert(sizeof(PixConst) == 192, "PixConst is 192 bytes");
// Executed by ONE warp (no barrier: every entry is written by the lane that loads it).  `tabs`: the image's W + H
// table entries.
template <int SHADER>
__device__ __forceinline__ void pix_const_write(const JrRenderArgs& a, int b, PixConst* __restrict__ c,
                                                float* __restrict__ tabs) {
  const int t = threadIdx.x & 31;
  {
    const float* __restrict__ vp = a.viewport.ptr + (long long)b * a.viewport.batch_stride;
    const float vp0 = vp[0], vp3 = vp[3], vp5 = vp[5], vp7 = vp[7];
    for (int i = t; i < a.W; i += 32) tabs[i] = fdiv_z((float)i - vp3, vp0);          // frag_setup_pix's expressions
    for (int i = t; i < a.H; i += 32) tabs[a.W + i] = fdiv_z((float)i - vp7, vp5);
    if (t == 11) { c->xs = tabs; c->ys = tabs + a.W; }
  }
  if (t < 6) {
    const int k = t == 0 ? 0 : (t == 1 ? 3 : (t == 2 ? 5 : (t == 3 ? 7 : (t == 4 ? 10 : 11))));
    reinterpret_cast<float*>(c)[t] = (a.viewport.ptr + (long long)b * a.viewport.batch_stride)[k];
  } else if (t < 9) {
    if (SHADER != JR_DEPTH) c->lcol[t - 6] = (a.light_colour.ptr + (long long)b * a.light_colour.batch_stride)[t - 6];
  } else if (t == 9) {
    if (SHADER >= JR_PHONG) {
      const JrF32& ld = SHADER >= JR_PHONG_REFLECTION ? a.light_dir_eye : a.light_direction;
      float l[3];
      load3(ld, b, l);
      const Vec3 n = normalise3(Vec3{l[0], l[1], l[2]});
      c->nl[0] = n.x; c->nl[1] = n.y; c->nl[2] = n.z;
    }
  } else if (t == 10) {
    // m = floor(2^32 / H) + 1 = (2^32 + e) / H with 0 < e <= H: umulhi(p, m) == floor(p / H) while p * e < 2^32,
    // guaranteed for p < W * H when W * H * H < 2^32 (H >= 2: for H == 1 the magic number does not fit 32 bits)
    const unsigned long long whh = (unsigned long long)a.W * a.H * a.H;
    c->h_magic = (a.H >= 2 && whh < (1ull << 32)) ? (0xFFFFFFFFu / (unsigned)a.H + 1u) : 0u;
  } else if (t >= 12 && t < 15) {
    if (SHADER >= JR_PHONG_REFLECTION) {
      c->amb[t - 12] = (a.ambient.ptr + (long long)b * a.ambient.batch_stride)[t - 12];
      c->dif[t - 12] = (a.diffuse.ptr + (long long)b * a.diffuse.batch_stride)[t - 12];
      c->spe[t - 12] = (a.specular.ptr + (long long)b * a.specular.batch_stride)[t - 12];
    }
    if (SHADER == JR_PHONG_REFLECTION_SHADOW)
      c->str[t - 12] = (a.shadow_strength.ptr + (long long)b * a.shadow_strength.batch_stride)[t - 12];
  } else if (t >= 16) {
    if (SHADER == JR_PHONG_REFLECTION_SHADOW)
      c->svp[t - 16] = (a.shadow_viewport.ptr + (long long)b * a.shadow_viewport.batch_stride)[t - 16];
  }
}

// Per-PIXEL part: NDC position, edge functions, 1/w, depth, perspective-correct weights.
// Needs f.inv and f.cl[k][2].
template <bool PC = false>
__device__ __forceinline__ void frag_setup_pix(const JrRenderArgs& a, int b, int x, int y, Frag& f,
                                               const PixConst* __restrict__ pc = nullptr) {
  const float* __restrict__ vp = a.viewport.ptr + (long long)b * a.viewport.batch_stride;
  const float vp0 = PC ? pc->vp0 : vp[0], vp3 = PC ? pc->vp3 : vp[3], vp5 = PC ? pc->vp5 : vp[5], vp7 = PC ? pc->vp7 : vp[7];
  if (PC) {
    f.xn = pc->xs[x]; f.yn = pc->ys[y];
  } else {
    f.xn = fdiv_z((float)x - vp3, vp0);   // (zero at the centre column / row of the canvas)
    f.yn = fdiv_z((float)y - vp7, vp5);
  }
  clip_coef(f.inv, f.xn, f.yn, f.cc);
  f.w_rec = (f.cc[0] + f.cc[1]) + f.cc[2];
  f.z = (f.cc[0] * f.cl[0][2] + f.cc[1] * f.cl[1][2]) + f.cc[2] * f.cl[2][2];
  f.zw = f.z * (PC ? pc->vp10 : vp[10]) + (PC ? pc->vp11 : vp[11]);
#pragma unroll
  for (int k = 0; k < 3; ++k) f.tc[k] = f.cc[k] / f.w_rec;
}

// ---------------------------------------------------------------------------------------------
// Vertex stage of the chosen triangle (reference `Shader.vertex` of each built-in, evaluated for
// the triangle's three vertices): everything that does not depend on the pixel.
template <int SHADER>
__device__ __forceinline__ void frag_vertex(const JrRenderArgs& a, int b, int tri, Frag& f) {
  frag_setup_tri(a, b, tri, f);
  if constexpr (SHADER == JR_DEPTH) return;
#pragma unroll
  for (int k = 0; k < 3; ++k) { f.fn[k] = f.fi[k]; f.fu[k] = f.fi[k]; }
  if (a.faces_norm.ptr) {
    const int32_t* q = a.faces_norm.ptr + (long long)b * a.faces_norm.batch_stride + 3 * tri;
    f.fn[0] = q[0]; f.fn[1] = q[1]; f.fn[2] = q[2];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) f.fn[k] = min(max(f.fn[k], 0), a.n_nrm - 1);
  if (a.faces_uv.ptr) {
    const int32_t* q = a.faces_uv.ptr + (long long)b * a.faces_uv.batch_stride + 3 * tri;
    f.fu[0] = q[0]; f.fu[1] = q[1]; f.fu[2] = q[2];
  }
  if (SHADER >= JR_GOURAUD_TEXTURE) {
#pragma unroll
    for (int k = 0; k < 3; ++k) f.fu[k] = min(max(f.fu[k], 0), a.n_uv - 1);
  }
  const float* __restrict__ nrm = a.normal.ptr + (long long)b * a.normal.batch_stride;
  load3(a.light_colour, b, f.lcol);
#pragma unroll
  for (int k = 0; k < 3; ++k) f.nraw[k] = fetch_normal(a, b, nrm, f.fn[k]);
  if (SHADER >= JR_GOURAUD_TEXTURE) {
    const float* __restrict__ uvp = a.uv.ptr + (long long)b * a.uv.batch_stride;
#pragma unroll
    for (int k = 0; k < 3; ++k) { f.uvv[k][0] = uvp[2 * f.fu[k]]; f.uvv[k][1] = uvp[2 * f.fu[k] + 1]; }
  }

  if (SHADER == JR_GOURAUD || SHADER == JR_GOURAUD_TEXTURE) {
    load3(a.light_direction, b, f.lraw);
    f.nl = normalise3(Vec3{f.lraw[0], f.lraw[1], f.lraw[2]});
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      f.nvert[k] = normalise3(f.nraw[k]);
      f.inten[k] = dot3(f.nvert[k].x, f.nvert[k].y, f.nvert[k].z, f.nl.x, f.nl.y, f.nl.z);
    }
    if (SHADER == JR_GOURAUD) {
      const float* __restrict__ cv = a.colour.ptr + (long long)b * a.colour.batch_stride;
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) f.colv[k][c] = (cv[3 * f.fi[k] + c] * f.lcol[c]) * f.inten[k];
    }
    return;
  }

  // ---- S4-S7: eye-space normals (phong.py:92-103)
  const float* __restrict__ wen = a.world_to_eye_norm.ptr + (long long)b * a.world_to_eye_norm.batch_stride;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    // Camera.apply_vec(normalise(n), wen): normalise twice, rotate, normalise
    const Vec3 m = normalise3(normalise3(f.nraw[k]));
    Vec3 t;
    t.x = (m.x * wen[0] + m.y * wen[1]) + m.z * wen[2];
    t.y = (m.x * wen[4] + m.y * wen[5]) + m.z * wen[6];
    t.z = (m.x * wen[8] + m.y * wen[9]) + m.z * wen[10];
    f.mvert[k] = m; f.tvert[k] = t;
    f.nvert[k] = normalise3(t);
  }
  if (SHADER >= JR_PHONG_REFLECTION) {
    const int32_t* __restrict__ ftp =
        a.faces_tex.ptr ? a.faces_tex.ptr + (long long)b * a.faces_tex.batch_stride + 3 * tri : nullptr;
    const int tv = min(max(ftp ? ftp[0] : f.fi[0], 0), a.n_texidx - 1);
    f.ti = min(max((a.texture_index.ptr + (long long)b * a.texture_index.batch_stride)[tv], 0), a.n_objects - 1);
  }
  if (SHADER == JR_PHONG_REFLECTION_SHADOW) {
    const float* __restrict__ sw2c = a.shadow_world_to_clip.ptr + (long long)b * a.shadow_world_to_clip.batch_stride;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float sc[4];
      to_clip(sw2c, f.P[k].x, f.P[k].y, f.P[k].z, sc);
#pragma unroll
      for (int j = 0; j < 4; ++j) f.scv[k][j] = sc[j] / sc[3];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Interpolate + fragment + mix for pixel (x, y) given the triangle's vertex-stage outputs in `f`
// (either just computed by frag_vertex or loaded from a TriAttr record).  `tri` is only used by
// the Darboux shader.
template <int SHADER, bool PC = false>
__device__ __forceinline__ void frag_pixel(const JrRenderArgs& a, int b, int x, int y, Frag& f,
                                           const PixConst* __restrict__ pc = nullptr) {
  frag_setup_pix<PC>(a, b, x, y, f, pc);
  if (PC && SHADER != JR_DEPTH) { f.lcol[0] = pc->lcol[0]; f.lcol[1] = pc->lcol[1]; f.lcol[2] = pc->lcol[2]; }
  const float* tc = f.tc;
  f.col[0] = f.col[1] = f.col[2] = 0.f;
  f.keep = true;
  f.texel = -1;
  f.spec_idx = -1;
  if (SHADER == JR_DEPTH) return;

  if (SHADER == JR_GOURAUD) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      f.col[c] = interp3(tc, f.colv[0][c], f.colv[1][c], f.colv[2][c]);
      f.keep = f.keep && (f.col[c] >= 0.f);
    }
    return;
  }
  const float* __restrict__ tex = a.texture.ptr + (long long)b * a.texture.batch_stride;
  const float u = interp3(tc, f.uvv[0][0], f.uvv[1][0], f.uvv[2][0]);
  const float v = interp3(tc, f.uvv[0][1], f.uvv[1][1], f.uvv[2][1]);
  if (SHADER == JR_GOURAUD_TEXTURE) {
    const int ui = pymod((int)floorf(u), a.tex_w), vi = pymod((int)floorf(v), a.tex_h);
    f.texel = (long long)ui * a.tex_h + vi;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      f.tex[c] = tex[f.texel * 3 + c];
      f.lc[c] = interp3(tc, f.lcol[c] * f.inten[0], f.lcol[c] * f.inten[1], f.lcol[c] * f.inten[2]);
      f.keep = f.keep && (f.lc[c] >= 0.f);
      f.col[c] = f.tex[c] * f.lc[c];
    }
    return;
  }

  f.normal = Vec3{interp3(tc, f.nvert[0].x, f.nvert[1].x, f.nvert[2].x),
                  interp3(tc, f.nvert[0].y, f.nvert[1].y, f.nvert[2].y),
                  interp3(tc, f.nvert[0].z, f.nvert[1].z, f.nvert[2].z)};
  // continuous lighting only -- except for the Darboux shader, whose tangent frame inverts a matrix built on this
  // normal (ill-conditioned: a last-bit change of nn showed up as 4e-5 in the colour): exact there
  f.nn = SHADER == JR_PHONG_DARBOUX ? normalise3(f.normal) : normalise3_fast(f.normal);

  if (SHADER == JR_PHONG || SHADER == JR_PHONG_DARBOUX) {
    if (PC) {
      f.nl = Vec3{pc->nl[0], pc->nl[1], pc->nl[2]};
    } else {
      load3(a.light_direction, b, f.lraw);
      f.nl = normalise3(Vec3{f.lraw[0], f.lraw[1], f.lraw[2]});
    }
    const int ui = pymod((int)floorf(u), a.tex_w), vi = pymod((int)floorf(v), a.tex_h);
    f.texel = (long long)ui * a.tex_h + vi;
    Vec3 nn = f.nn;
    if (SHADER == JR_PHONG_DARBOUX) {
      // phong_darboux.py:144-151, :231-262
      const float* __restrict__ w2c = a.world_to_clip.ptr + (long long)b * a.world_to_clip.batch_stride;
      const float* __restrict__ pos = a.position.ptr + (long long)b * a.position.batch_stride;
      const float* __restrict__ uvp = a.uv.ptr + (long long)b * a.uv.batch_stride;
      const int32_t* __restrict__ i2f = a.id_to_face.ptr + (long long)b * a.id_to_face.batch_stride;
      const int32_t* __restrict__ fidx = a.faces_indices.ptr + (long long)b * a.faces_indices.batch_stride;
      // gathers wrap one negative and clamp, like the reference's jnp indexing (and never leave the arrays)
      const int face = wrap_clamp(i2f[f.fi[0]], a.n_faces_indices);
      float tr[3][3], tuv[3][2];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int vraw = fidx[3 * face + k];
        const int vtx = wrap_clamp(vraw, a.n_pos);
        const int vuv = wrap_clamp(vraw, a.n_uv);
        float tcq[4];
        to_clip(w2c, pos[3 * vtx], pos[3 * vtx + 1], pos[3 * vtx + 2], tcq);
        const bool w0 = tcq[3] == 0.0f;
        tr[k][0] = w0 ? tcq[0] : tcq[0] / tcq[3];
        tr[k][1] = w0 ? tcq[1] : tcq[1] / tcq[3];
        tr[k][2] = w0 ? tcq[2] : tcq[2] / tcq[3];
        tuv[k][0] = uvp[2 * vuv]; tuv[k][1] = uvp[2 * vuv + 1];
        f.dvtx[k] = vtx;
        f.dvuv[k] = vuv;
        f.dP[k] = Vec3{pos[3 * vtx], pos[3 * vtx + 1], pos[3 * vtx + 2]};
        f.dtr[k][0] = tr[k][0]; f.dtr[k][1] = tr[k][1]; f.dtr[k][2] = tr[k][2];
        f.dtw[k] = tcq[3];
      }
      const float A[9] = {tr[1][0] - tr[0][0], tr[1][1] - tr[0][1], tr[1][2] - tr[0][2],
                          tr[2][0] - tr[0][0], tr[2][1] - tr[0][1], tr[2][2] - tr[0][2],
                          nn.x, nn.y, nn.z};
      float AI[9];
      lu_inverse3(A, AI);
      const float du0 = tuv[1][0] - tuv[0][0], du1 = tuv[2][0] - tuv[0][0];
      const float dv0 = tuv[1][1] - tuv[0][1], dv1 = tuv[2][1] - tuv[0][1];
      const Vec3 ivr = {AI[0] * du0 + AI[1] * du1, AI[3] * du0 + AI[4] * du1, AI[6] * du0 + AI[7] * du1};
      const Vec3 jvr = {AI[0] * dv0 + AI[1] * dv1, AI[3] * dv0 + AI[4] * dv1, AI[6] * dv0 + AI[7] * dv1};
      const Vec3 iv = normalise3(ivr);
      const Vec3 jv = normalise3(jvr);
      const float* nm = a.normal_map.ptr + (long long)b * a.normal_map.batch_stride + f.texel * 3;
      const Vec3 bn = {(iv.x * nm[0] + jv.x * nm[1]) + nn.x * nm[2],
                       (iv.y * nm[0] + jv.y * nm[1]) + nn.y * nm[2],
                       (iv.z * nm[0] + jv.z * nm[1]) + nn.z * nm[2]};
      nn = normalise3(bn);
#pragma unroll
      for (int k = 0; k < 9; ++k) f.dAI[k] = AI[k];
      f.dduv[0] = du0; f.dduv[1] = du1; f.dduv[2] = dv0; f.dduv[3] = dv1;
      f.divr = ivr; f.djvr = jvr; f.div_ = iv; f.djv = jv;
      f.dnm[0] = nm[0]; f.dnm[1] = nm[1]; f.dnm[2] = nm[2];
      f.dbn = bn; f.dn2 = nn;
    }
    f.ndl = dot3(nn.x, nn.y, nn.z, f.nl.x, f.nl.y, f.nl.z);
    f.ok = true;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      f.tex[c] = tex[f.texel * 3 + c];
      f.lc[c] = f.lcol[c] * f.ndl;
      f.ok = f.ok && (f.lc[c] >= 0.f);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) f.col[c] = f.ok ? f.tex[c] * f.lc[c] : 0.f;
    return;
  }

  // ---- S6 / S7 (phong_reflection.py:175-220, phong_reflection_shadow.py:196-257)
  const int ti = f.ti;
  const int32_t* tsh = a.texture_shape.ptr + (long long)b * a.texture_shape.batch_stride + 2 * ti;
  float fu0 = u - truncf(u), fv0 = v - truncf(v);  // jnp.modf(uv)[0]
  if (fu0 < 0.f) fu0 = fu0 + 1.f;
  if (fv0 < 0.f) fv0 = fv0 + 1.f;
  const float ur = fu0 * (float)tsh[0] + (float)(ti * a.texture_offset);
  const float vr = fv0 * (float)tsh[1];
  const int U = (int)floorf(ur), V = (int)floorf(vr);
  f.texel = (long long)wrap_clamp(U, a.tex_w) * a.tex_h + wrap_clamp(V, a.tex_h);
  f.spec_idx = (long long)wrap_clamp(U, a.spec_w) * a.spec_h + wrap_clamp(V, a.spec_h);
  if (PC) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { f.amb[c] = pc->amb[c]; f.dif[c] = pc->dif[c]; f.spe[c] = pc->spe[c]; }
    f.nl = Vec3{pc->nl[0], pc->nl[1], pc->nl[2]};
  } else {
    load3(a.light_dir_eye, b, f.lraw);
    load3(a.ambient, b, f.amb);
    load3(a.diffuse, b, f.dif);
    load3(a.specular, b, f.spe);
    f.nl = normalise3(Vec3{f.lraw[0], f.lraw[1], f.lraw[2]});
  }
  const Vec3 nn = f.nn, ld = f.nl;
  f.ndl = dot3(nn.x, nn.y, nn.z, ld.x, ld.y, ld.z);
  f.diffuse = fmaxf(f.ndl, 0.f);
  const float two_ndl = 2.f * f.ndl;
  f.rv = Vec3{two_ndl * nn.x - ld.x, two_ndl * nn.y - ld.y, two_ndl * nn.z - ld.z};
  const Vec3 refl = normalise3_fast(f.rv);
  f.sexp = (a.specular_map.ptr + (long long)b * a.specular_map.batch_stride)[f.spec_idx];
  f.base = fmaxf(refl.z, 0.f);
  f.specular = pow_unit(f.base, f.sexp);
  f.shadow[0] = f.shadow[1] = f.shadow[2] = 1.f;
  f.lit = true;
  if (SHADER == JR_PHONG_REFLECTION_SHADOW) {
    const float* __restrict__ svp = PC ? pc->svp : a.shadow_viewport.ptr + (long long)b * a.shadow_viewport.batch_stride;
    float sc[4], ss[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) sc[j] = interp3(tc, f.scv[0][j], f.scv[1][j], f.scv[2][j]);
#pragma unroll
    for (int r = 0; r < 4; ++r)
      ss[r] = ((svp[4 * r] * sc[0] + svp[4 * r + 1] * sc[1]) + svp[4 * r + 2] * sc[2]) + svp[4 * r + 3] * sc[3];
    const float sx = ss[0] / ss[3], sy = ss[1] / ss[3], sz = ss[2] / ss[3];
    // Shadow.get (shadow.py:129-153): round half away, one negative wrap, OOB -> +inf
    float rx = roundf(sx), ry = roundf(sy);
    const bool finite = (rx == rx) && (ry == ry);
    rx = fminf(fmaxf(rx, -1e9f), 1e9f);
    ry = fminf(fmaxf(ry, -1e9f), 1e9f);
    int px = (int)rx, py = (int)ry;
    if (px < 0) px += a.shadow_w;
    if (py < 0) py += a.shadow_h;
    float sval = __int_as_float(0x7f800000);
    if (finite && px >= 0 && px < a.shadow_w && py >= 0 && py < a.shadow_h)
      sval = (a.shadow_map.ptr + (long long)b * a.shadow_map.batch_stride)[(long long)px * a.shadow_h + py];
    f.lit = sz <= sval;
    if (PC) { f.str[0] = pc->str[0]; f.str[1] = pc->str[1]; f.str[2] = pc->str[2]; }
    else load3(a.shadow_strength, b, f.str);
#pragma unroll
    for (int c = 0; c < 3; ++c) f.shadow[c] = f.lit ? 1.f : 1.f - f.str[c];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    f.tex[c] = tex[f.texel * 3 + c];
    f.ds[c] = f.dif[c] * f.diffuse + f.spe[c] * f.specular;
    if (SHADER == JR_PHONG_REFLECTION)
      f.col[c] = f.amb[c] * f.tex[c] + (f.ds[c] * f.lcol[c]) * f.tex[c];
    else
      f.col[c] = f.amb[c] * f.tex[c] + ((f.shadow[c] * f.ds[c]) * f.tex[c]) * f.lcol[c];
  }
}

// Both stages for one pixel (backward and small-canvas forward: recompute per pixel).
template <int SHADER>
__device__ __forceinline__ void shade_pixel(const JrRenderArgs& a, int b, int x, int y, int tri, Frag& f) {
  frag_vertex<SHADER>(a, b, tri, f);
  frag_pixel<SHADER>(a, b, x, y, f);
}

// ---------------------------------------------------------------------------------------------
// Per-triangle attribute record: the vertex-stage outputs the pixel stage reads, written once per
// triangle by k_tri_attr and shared by all pixels of the triangle (large canvases).
constexpr int TA_FLOATS = 44;  // 176 bytes, 11 x float4
// layout: [0..8] inv, [9..11] zc, [12..20] nvert (S4-S7) or colv (S2), [21..26] uvv, [27] ti,
//         [28..39] scv (S7), [40..42] inten (S3), [43] pad
template <int SHADER>
__device__ __forceinline__ void attr_store(const Frag& f, float* __restrict__ r) {
#pragma unroll
  for (int k = 0; k < 9; ++k) r[k] = f.inv[k];
  r[9] = f.cl[0][2]; r[10] = f.cl[1][2]; r[11] = f.cl[2][2];
  if (SHADER == JR_GOURAUD) {
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) r[12 + 3 * k + c] = f.colv[k][c];
  }
  if (SHADER >= JR_PHONG) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { r[12 + 3 * k] = f.nvert[k].x; r[13 + 3 * k] = f.nvert[k].y; r[14 + 3 * k] = f.nvert[k].z; }
  }
  if (SHADER >= JR_GOURAUD_TEXTURE) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { r[21 + 2 * k] = f.uvv[k][0]; r[22 + 2 * k] = f.uvv[k][1]; }
  }
  if (SHADER >= JR_PHONG_REFLECTION) r[27] = __int_as_float(f.ti);
  if (SHADER == JR_PHONG_REFLECTION_SHADOW) {
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) r[28 + 4 * k + j] = f.scv[k][j];
  }
  if (SHADER == JR_GOURAUD_TEXTURE) { r[40] = f.inten[0]; r[41] = f.inten[1]; r[42] = f.inten[2]; }
}

template <int SHADER, bool PC = false>
__device__ __forceinline__ void attr_load(const JrRenderArgs& a, int b, const float* __restrict__ r, Frag& f) {
#pragma unroll
  for (int k = 0; k < 9; ++k) f.inv[k] = r[k];
  f.cl[0][2] = r[9]; f.cl[1][2] = r[10]; f.cl[2][2] = r[11];
  if (SHADER == JR_DEPTH) return;
  if (!PC) load3(a.light_colour, b, f.lcol);   // (PC: frag_pixel takes it from the staged constants)
  if (SHADER == JR_GOURAUD) {
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) f.colv[k][c] = r[12 + 3 * k + c];
  }
  if (SHADER >= JR_PHONG) {
#pragma unroll
    for (int k = 0; k < 3; ++k) f.nvert[k] = Vec3{r[12 + 3 * k], r[13 + 3 * k], r[14 + 3 * k]};
  }
  if (SHADER >= JR_GOURAUD_TEXTURE) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { f.uvv[k][0] = r[21 + 2 * k]; f.uvv[k][1] = r[22 + 2 * k]; }
  }
  if (SHADER >= JR_PHONG_REFLECTION) f.ti = __float_as_int(r[27]);
  if (SHADER == JR_PHONG_REFLECTION_SHADOW) {
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) f.scv[k][j] = r[28 + 4 * k + j];
  }
  if (SHADER == JR_GOURAUD_TEXTURE) { f.inten[0] = r[40]; f.inten[1] = r[41]; f.inten[2] = r[42]; }
}

// Visible-triangle list: only triangles that won at least one pixel get an attribute record (at 84x84
// about one Brax triangle in eight; writing all records was the one HBM-bound kernel of the facade).
// The first pixel to flag a triangle (atomicOr on the flag's 32-bit word) appends it to the image's
// list; the order of the list is irrelevant, every record goes to its own slot.
template <int UNUSED>  // template only for inline linkage (two translation units include this header)
__global__ void __launch_bounds__(256) k_mark_visible(const int32_t* __restrict__ tri_id, unsigned* __restrict__ flag_words,
                                                      int* __restrict__ list, int* __restrict__ count, int npix, int T,
                                                      int B, int* __restrict__ slot_map = nullptr) {
  for (int b = blockIdx.y; b < B; b += gridDim.y)
    for (int p0 = blockIdx.x * 256; p0 < npix; p0 += gridDim.x * 256) {
      const int pix = p0 + threadIdx.x;
      const int tri = pix < npix ? tri_id[(long long)b * npix + pix] : -1;
      // neighbouring pixels mostly share their triangle: one lane per run of equal ids goes on
      const int prev = __shfl_up_sync(0xffffffffu, tri, 1);
      if (tri < 0 || ((threadIdx.x & 31) != 0 && prev == tri)) continue;
      const long long bit = (long long)b * T + tri;
      unsigned* w = flag_words + (bit >> 5);
      const unsigned m = 1u << (bit & 31);
      if (*w & m) continue;                    // already listed (plain load first: most pixels stop here)
      if (atomicOr(w, m) & m) continue;
      const int slot = atomicAdd(&count[b], 1);
      list[(long long)b * T + slot] = tri;
      if (slot_map) slot_map[(long long)b * T + tri] = slot;  // compact records: triangle -> record slot
    }
}

// ---------------------------------------------------------------------------------------------
// Extended record for the backward pass of the phong_reflection* shaders: the attribute record plus the
// vertex-stage intermediates that only the reverse pass reads (world positions, raw / rotated normals).
constexpr int TE_FLOATS = 80;  // 320 bytes, 20 x float4: [0..43] attribute record, [44..52] P, [53..61] nraw,
                               // [62..70] mvert, [71..79] tvert
template <int SHADER>
__device__ __forceinline__ void rec_store(const Frag& f, float* __restrict__ r) {
  attr_store<SHADER>(f, r);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    r[44 + 3 * k] = f.P[k].x; r[45 + 3 * k] = f.P[k].y; r[46 + 3 * k] = f.P[k].z;
    r[53 + 3 * k] = f.nraw[k].x; r[54 + 3 * k] = f.nraw[k].y; r[55 + 3 * k] = f.nraw[k].z;
    r[62 + 3 * k] = f.mvert[k].x; r[63 + 3 * k] = f.mvert[k].y; r[64 + 3 * k] = f.mvert[k].z;
    r[71 + 3 * k] = f.tvert[k].x; r[72 + 3 * k] = f.tvert[k].y; r[73 + 3 * k] = f.tvert[k].z;
  }
}
template <int SHADER>
__device__ __forceinline__ void rec_load(const JrRenderArgs& a, int b, const float* __restrict__ r, Frag& f) {
  attr_load<SHADER>(a, b, r, f);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    f.P[k] = Vec3{r[44 + 3 * k], r[45 + 3 * k], r[46 + 3 * k]};
    f.nraw[k] = Vec3{r[53 + 3 * k], r[54 + 3 * k], r[55 + 3 * k]};
    f.mvert[k] = Vec3{r[62 + 3 * k], r[63 + 3 * k], r[64 + 3 * k]};
    f.tvert[k] = Vec3{r[71 + 3 * k], r[72 + 3 * k], r[73 + 3 * k]};
  }
}

}  // namespace jr
