// jr_geometry.cuh -- vertex / normal fetch shared by every forward kernel: plain loads of the merged world-space
// arrays, or -- instanced geometry (JrRenderArgs.inst_*) -- the world-space merge of
// `merge_objects` (renderer/model.py:447-555) evaluated on the fly from shared local meshes and per-image object
// transforms, in the arithmetic of k_merge_verts / k_merge_norms (jr_forward.cu): same bits as the merged path.
#pragma once
#include "../../include/jr_b200.h"
#include "jr_device.cuh"

namespace jr {

__device__ __forceinline__ bool instanced(const JrRenderArgs& a) { return a.inst_transform.ptr != nullptr; }

// to_cartesian (geometry.py:183-202) of a homogeneous point: h / h.w unless w == 0.  w == 1 -- every affine object
// transform -- skips the three IEEE divisions: x / 1 == x bit for bit.
__device__ __forceinline__ void to_cartesian3(const float h[4], float& x, float& y, float& z) {
  const bool keep = h[3] == 0.0f || h[3] == 1.0f;
  x = keep ? h[0] : h[0] / h[3];
  y = keep ? h[1] : h[1] / h[3];
  z = keep ? h[2] : h[2] / h[3];
}

// local vertex (lx, ly, lz) with index i of image b -> world space (model.py:489-499; geometry.py:183-202)
__device__ __forceinline__ void instance_vertex(const JrRenderArgs& a, int b, int i, float lx, float ly, float lz,
                                                float& x, float& y, float& z) {
  const int o = min(max((a.inst_vert_object.ptr + (long long)b * a.inst_vert_object.batch_stride)[i], 0), a.n_inst - 1);
  const float* __restrict__ s = a.inst_scaling.ptr + (long long)b * a.inst_scaling.batch_stride + 3 * o;
  const float* __restrict__ T = a.inst_transform.ptr + (long long)b * a.inst_transform.batch_stride + 16 * o;
  float h[4];
  to_clip(T, lx * s[0], ly * s[1], lz * s[2], h);  // to_homogeneous(p * scaling) @ T^T
  to_cartesian3(h, x, y, z);
}

// The three vertices of one triangle (p = x0 y0 z0 x1 y1 z1 x2 y2 z2, local -> world in place).  A triangle's
// vertices normally belong to ONE object: its scaling and transform are then fetched once (19 loads instead of 57;
// the visibility kernels do this for every triangle of every image).  Same arithmetic as instance_vertex.
__device__ __forceinline__ void instance_triangle(const JrRenderArgs& a, int b, int i0, int i1, int i2, float p[9]) {
  const int32_t* __restrict__ vo = a.inst_vert_object.ptr + (long long)b * a.inst_vert_object.batch_stride;
  const int o0 = min(max(vo[i0], 0), a.n_inst - 1), o1 = min(max(vo[i1], 0), a.n_inst - 1), o2 = min(max(vo[i2], 0), a.n_inst - 1);
  if (o0 == o1 && o0 == o2) {
    const float* __restrict__ s = a.inst_scaling.ptr + (long long)b * a.inst_scaling.batch_stride + 3 * o0;
    const float* __restrict__ T = a.inst_transform.ptr + (long long)b * a.inst_transform.batch_stride + 16 * o0;
    const float s0 = s[0], s1 = s[1], s2 = s[2];
    float m[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m[k] = T[k];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      float h[4];
      to_clip(m, p[3 * v] * s0, p[3 * v + 1] * s1, p[3 * v + 2] * s2, h);
      to_cartesian3(h, p[3 * v], p[3 * v + 1], p[3 * v + 2]);
    }
  } else {
    instance_vertex(a, b, i0, p[0], p[1], p[2], p[0], p[1], p[2]);
    instance_vertex(a, b, i1, p[3], p[4], p[5], p[3], p[4], p[5]);
    instance_vertex(a, b, i2, p[6], p[7], p[8], p[6], p[7], p[8]);
  }
}

// world-space position of vertex i (pos_b = a.position.ptr + b * batch_stride)
__device__ __forceinline__ Vec3 fetch_position(const JrRenderArgs& a, int b, const float* __restrict__ pos_b, int i) {
  Vec3 v{pos_b[3 * i], pos_b[3 * i + 1], pos_b[3 * i + 2]};
  if (instanced(a)) instance_vertex(a, b, i, v.x, v.y, v.z, v.x, v.y, v.z);
  return v;
}

// world-space normal j (nrm_b = a.normal.ptr + b * batch_stride): model.py:517-530 as k_merge_norms computes it
__device__ __forceinline__ Vec3 fetch_normal(const JrRenderArgs& a, int b, const float* __restrict__ nrm_b, int j) {
  Vec3 n{nrm_b[3 * j], nrm_b[3 * j + 1], nrm_b[3 * j + 2]};
  if (instanced(a)) {
    const int o = min(max((a.inst_norm_object.ptr + (long long)b * a.inst_norm_object.batch_stride)[j], 0), a.n_inst - 1);
    const float* __restrict__ R = a.inst_normal_matrix.ptr + (long long)b * a.inst_normal_matrix.batch_stride + 16 * o;
    const float* __restrict__ f = a.inst_norm_scale.ptr + (long long)b * a.inst_norm_scale.batch_stride + 2 * o;
    const float f1 = f[0], f2 = f[1];
    // (fdiv_z: local normals are full of exact zeros -- cube faces, capsule poles)
    const float x = fdiv_z(n.x, f1), y = fdiv_z(n.y, f1), z = fdiv_z(n.z, f1);
    n.x = fdiv_z((x * R[0] + y * R[1]) + z * R[2], f2);
    n.y = fdiv_z((x * R[4] + y * R[5]) + z * R[6], f2);
    n.z = fdiv_z((x * R[8] + y * R[9]) + z * R[10], f2);
  }
  return n;
}

// Exact (reference-order, no-FMA) per-triangle cull and conservative pixel bbox, shared by k_vis3's exact phase,
// the binned setup and the audit kernel: clip x / y / w of the three vertices (rows 0, 1, 3 of world_to_clip) into
// M, `keep & front-facing` (det > 1e-6, pipeline.py:98-100, :232), "all w <= 0" (behind the camera), and -- when all
// w > 0 -- the screen bbox grown by bbox_margin (jr_device.cuh); any w <= 0: the whole canvas.  Returns false when
// the triangle cannot cover a sample.  Bbox inclusive, in pixels of a W x H canvas.
__device__ __forceinline__ bool exact_cull_bbox(const float* __restrict__ w2c, float vp00, float vp03, float vp11,
                                                float vp13, int W, int H, Vec3 q0, Vec3 q1, Vec3 q2, float M[9],
                                                int& x0, int& x1, int& y0, int& y1) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int rr = (r == 2) ? 3 : r;
    const float m0 = w2c[4 * rr], m1 = w2c[4 * rr + 1], m2 = w2c[4 * rr + 2], m3 = w2c[4 * rr + 3];
    M[0 + r] = ((q0.x * m0 + q0.y * m1) + q0.z * m2) + m3;
    M[3 + r] = ((q1.x * m0 + q1.y * m1) + q1.z * m2) + m3;
    M[6 + r] = ((q2.x * m0 + q2.y * m1) + q2.z * m2) + m3;
  }
  const float det = det3(M);
  const bool cand = det > 1e-6f;
  const float w0 = M[2], w1 = M[5], w2 = M[8];
  const bool behind = (w0 <= 0.f && w1 <= 0.f && w2 <= 0.f);
  if (!cand || behind) return false;
  x0 = 0; x1 = W - 1; y0 = 0; y1 = H - 1;
  if (w0 > 0.f && w1 > 0.f && w2 > 0.f) {
    const float r0 = __fdividef(1.f, w0), r1 = __fdividef(1.f, w1), r2 = __fdividef(1.f, w2);
    const float sx0 = (M[0] * r0) * vp00 + vp03, sx1 = (M[3] * r1) * vp00 + vp03, sx2 = (M[6] * r2) * vp00 + vp03;
    const float sy0 = (M[1] * r0) * vp11 + vp13, sy1 = (M[4] * r1) * vp11 + vp13, sy2 = (M[7] * r2) * vp11 + vp13;
    const float mg = bbox_margin(sx0, sy0, sx1, sy1, sx2, sy2, 2.f * fmaxf(vp00, vp11));
    const float mnx = fmaxf(fminf(fminf(sx0, sx1), sx2) - mg, 0.f);
    const float mxx = fminf(fmaxf(fmaxf(sx0, sx1), sx2) + mg, (float)(W - 1));
    const float mny = fmaxf(fminf(fminf(sy0, sy1), sy2) - mg, 0.f);
    const float mxy = fminf(fmaxf(fmaxf(sy0, sy1), sy2) + mg, (float)(H - 1));
    if (!(mnx <= mxx) || !(mny <= mxy)) return false;
    x0 = (int)ceilf(mnx); x1 = (int)floorf(mxx);
    y0 = (int)ceilf(mny); y1 = (int)floorf(mxy);
    if (x0 > x1 || y0 > y1) return false;
  }
  return true;
}

}  // namespace jr
