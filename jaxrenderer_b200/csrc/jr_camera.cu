// jr_camera.cu -- fused camera construction (include/jr_b200.h: jr_camera_build).
//
// One thread builds the 8 matrices of one batch element.  The arithmetic follows the reference's
// formulas step by step (geometry.py:235-278, :472-511, :536-763, :813-845) in a fixed scalar fp32
// order (-fmad=false); 4x4 products accumulate k = 0..3 in order.  This replaces ~120 framework ops
// (each a kernel launch) per camera; it is data preparation, not part of the parity-critical path:
// the render kernels take whatever matrices they are given.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/jr_b200.h"
#include "jr_common.cuh"
#include "jr_device.cuh"

namespace jr {

__device__ __forceinline__ void mat4_zero(float* m) {
#pragma unroll
  for (int i = 0; i < 16; ++i) m[i] = 0.f;
}
__device__ __forceinline__ void mat4_mul(const float* A, const float* B, float* C) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      C[4 * i + j] = ((A[4 * i] * B[j] + A[4 * i + 1] * B[4 + j]) + A[4 * i + 2] * B[8 + j]) + A[4 * i + 3] * B[12 + j];
}
__device__ __forceinline__ Vec3 cross3(Vec3 a, Vec3 b) {
  return Vec3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Camera.inv_scale_translation_matrix (geometry.py:472-511)
__device__ __forceinline__ void inv_scale_translation(const float* m, float* out) {
  float r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = 1.0f / m[5 * i];
  mat4_zero(out);
#pragma unroll
  for (int i = 0; i < 4; ++i) out[5 * i] = r[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) out[4 * i + 3] = (-(r[i] * m[4 * i + 3])) * r[3];
}

__global__ void __launch_bounds__(128) k_camera(const __grid_constant__ JrCameraArgs a) {
  const int b = blockIdx.x * 128 + threadIdx.x;
  if (b >= a.B) return;
  const float* p = a.params.ptr + (long long)b * a.params.batch_stride;
  Vec3 eye, centre{p[3], p[4], p[5]}, up{p[6], p[7], p[8]};
  float proj[16], vp[16];
  mat4_zero(proj);
  if (a.mode == JR_CAMERA_PERSPECTIVE) {
    eye = Vec3{p[0], p[1], p[2]};
    const float kRad = 0.017453292519943295f;
    const float tv = tanf((p[9] * kRad) / 2.0f), th = tanf((p[10] * kRad) / 2.0f);
    const float f = 1.0f / tv, aspect = th / tv, zn = p[11], zf = p[12];
    proj[0] = f / aspect;
    proj[5] = f;
    proj[10] = (zf + zn) / (zn - zf);
    proj[11] = ((2.0f * zf) * zn) / (zn - zf);
    proj[14] = -1.0f;
    mat4_zero(vp);
    const float w = p[13], h = p[14], d = p[15];
    vp[0] = w / 2.0f; vp[3] = 0.0f + w / 2.0f;
    vp[5] = h / 2.0f; vp[7] = 0.0f + h / 2.0f;
    vp[10] = d / 2.0f; vp[11] = d / 2.0f;
    vp[15] = 1.0f;
  } else {
    // eye = centre + light_direction * distance (shadow.py:84-88); here p[0..2] is the centre
    centre = Vec3{p[0], p[1], p[2]};
    eye = Vec3{centre.x + p[3] * p[9], centre.y + p[4] * p[9], centre.z + p[5] * p[9]};
    const float l = p[10], r = p[11], bo = p[12], t = p[13], n = p[14], fa = p[15];
    proj[0] = 2.0f / (r - l);
    proj[5] = 2.0f / (t - bo);
    proj[10] = -2.0f / (fa - n);
    proj[15] = 1.0f;
    proj[3] = -(r + l) / (r - l);
    proj[7] = -(t + bo) / (t - bo);
    proj[11] = -(fa + n) / (fa - n);
    const float* v = a.viewport.ptr + (long long)b * a.viewport.batch_stride;
#pragma unroll
    for (int i = 0; i < 16; ++i) vp[i] = v[i];
  }
  // lookAt (geometry.py:536-575) and its analytic inverse (:577-636)
  const Vec3 fwd = normalise3(Vec3{centre.x - eye.x, centre.y - eye.y, centre.z - eye.z});
  const Vec3 upn = normalise3(up);
  const Vec3 side = normalise3(cross3(fwd, upn));
  const Vec3 up2 = cross3(side, fwd);
  const float R[9] = {side.x, side.y, side.z, up2.x, up2.y, up2.z, -fwd.x, -fwd.y, -fwd.z};
  float view[16], view_inv[16];
  mat4_zero(view); mat4_zero(view_inv);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) { view[4 * i + j] = R[3 * i + j]; view_inv[4 * j + i] = R[3 * i + j]; }
    view[4 * i + 3] = -((R[3 * i] * eye.x + R[3 * i + 1] * eye.y) + R[3 * i + 2] * eye.z);
  }
  view_inv[3] = eye.x; view_inv[7] = eye.y; view_inv[11] = eye.z;
  view[15] = 1.0f; view_inv[15] = 1.0f;

  // projection inverse: perspective (columns / rows 2 and 3 swapped around the scale-translation
  // inverse, geometry.py:686-718) when projection[3][3] is ~0, orthographic otherwise
  float proj_inv[16], tmp[16], sh[16];
  if (fabsf(proj[15]) <= 1e-8f) {
    const int s[4] = {0, 1, 3, 2};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sh[4 * i + j] = proj[4 * i + s[j]];
    inv_scale_translation(sh, tmp);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) proj_inv[4 * i + j] = tmp[4 * s[i] + j];
  } else {
    inv_scale_translation(proj, proj_inv);
  }
  float vp_inv[16];
  inv_scale_translation(vp, vp_inv);

  float w2c[16], w2s[16], s2w[16];
  mat4_mul(proj, view, w2c);
  mat4_mul(vp, proj, tmp);
  mat4_mul(tmp, view, w2s);
  mat4_mul(view_inv, proj_inv, tmp);
  mat4_mul(tmp, vp_inv, s2w);

  const long long plane = (long long)a.B * 16;
  float* o = a.out + (long long)b * 16;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    o[0 * plane + i] = view[i];
    o[1 * plane + i] = proj[i];
    o[2 * plane + i] = vp[i];
    o[3 * plane + i] = w2c[i];
    o[4 * plane + i] = view_inv[4 * (i & 3) + (i >> 2)];  // world_to_eye_norm = view_inv^T
    o[5 * plane + i] = w2s[i];
    o[6 * plane + i] = view_inv[i];
    o[7 * plane + i] = s2w[i];
  }
}

}  // namespace jr

extern "C" int jr_camera_build(const JrCameraArgs* a, jr_stream_t stream) {
  if (!a || !a->params.ptr || !a->out) return JR_ERR_NULL;
  if (a->B <= 0) return JR_ERR_DIMS;
  if (a->mode != JR_CAMERA_PERSPECTIVE && a->mode != JR_CAMERA_LIGHT) return JR_ERR_UNSUPPORTED;
  if (a->mode == JR_CAMERA_LIGHT && !a->viewport.ptr) return JR_ERR_NULL;
  jr::k_camera<<<(a->B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*a);
  jr::g_launches++;
  return cudaGetLastError() == cudaSuccess ? JR_OK : JR_ERR_CUDA;
}
