// jr_geometry.cuh -- vertex / normal fetch shared by every forward kernel: plain loads of the merged world-space
// arrays, or -- instanced geometry (JrRenderArgs.inst_*) -- the world-space merge of
// `merge_objects` (renderer/model.py:447-555) evaluated on the fly from shared local meshes and per-image object
// transforms, in the arithmetic of k_merge_verts / k_merge_norms (jr_forward.cu): same bits as the merged path.
#pragma once
#include "../../include/jr_b200.h"
#include "jr_device.cuh"

namespace jr {

__device__ __forceinline__ bool instanced(const JrRenderArgs& a) { return a.inst_transform.ptr != nullptr; }

// local vertex (lx, ly, lz) with index i of image b -> world space (model.py:489-499; geometry.py:183-202)
__device__ __forceinline__ void instance_vertex(const JrRenderArgs& a, int b, int i, float lx, float ly, float lz,
                                                float& x, float& y, float& z) {
  const int o = min(max((a.inst_vert_object.ptr + (long long)b * a.inst_vert_object.batch_stride)[i], 0), a.n_inst - 1);
  const float* __restrict__ s = a.inst_scaling.ptr + (long long)b * a.inst_scaling.batch_stride + 3 * o;
  const float* __restrict__ T = a.inst_transform.ptr + (long long)b * a.inst_transform.batch_stride + 16 * o;
  float h[4];
  to_clip(T, lx * s[0], ly * s[1], lz * s[2], h);  // to_homogeneous(p * scaling) @ T^T
  const bool w0 = h[3] == 0.0f;                      // to_cartesian
  x = w0 ? h[0] : h[0] / h[3];
  y = w0 ? h[1] : h[1] / h[3];
  z = w0 ? h[2] : h[2] / h[3];
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
    const float x = n.x / f1, y = n.y / f1, z = n.z / f1;
    n.x = ((x * R[0] + y * R[1]) + z * R[2]) / f2;
    n.y = ((x * R[4] + y * R[5]) + z * R[6]) / f2;
    n.z = ((x * R[8] + y * R[9]) + z * R[10]) / f2;
  }
  return n;
}

}  // namespace jr
