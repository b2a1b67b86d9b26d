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

// ---- scalar types the camera formulas are evaluated in: float (forward) and a dual number (value, derivative)
// for the reverse pass (jr_camera_vjp): the SAME templated formulas, so the derivative is that of the code above
// it, operation by operation.  For float every expression keeps the order of the original kernel (same bits).
struct Dual {
  float v, d;
  __device__ Dual() : v(0.f), d(0.f) {}
  __device__ Dual(float v_) : v(v_), d(0.f) {}
  __device__ Dual(float v_, float d_) : v(v_), d(d_) {}
};
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
__device__ __forceinline__ Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
  const float q = a.v / b.v;
  return Dual(q, (a.d - q * b.d) / b.v);
}
__device__ __forceinline__ float val(float x) { return x; }
__device__ __forceinline__ float val(Dual x) { return x.v; }
__device__ __forceinline__ float sqrt_t(float x) { return sqrtf(x); }
__device__ __forceinline__ Dual sqrt_t(Dual x) { const float s = sqrtf(x.v); return Dual(s, x.d / (2.0f * s)); }
__device__ __forceinline__ float tan_t(float x) { return tanf(x); }
__device__ __forceinline__ Dual tan_t(Dual x) { const float t = tanf(x.v); return Dual(t, x.d * (1.0f + t * t)); }

template <typename T> struct V3 { T x, y, z; };

template <typename T>
__device__ __forceinline__ void mat4_zero(T* m) {
#pragma unroll
  for (int i = 0; i < 16; ++i) m[i] = T(0.f);
}
template <typename T>
__device__ __forceinline__ void mat4_mul(const T* A, const T* B, T* C) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      C[4 * i + j] = ((A[4 * i] * B[j] + A[4 * i + 1] * B[4 + j]) + A[4 * i + 2] * B[8 + j]) + A[4 * i + 3] * B[12 + j];
}
template <typename T>
__device__ __forceinline__ V3<T> cross_t(V3<T> a, V3<T> b) {
  return V3<T>{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <typename T>
__device__ __forceinline__ V3<T> normalise_t(V3<T> v) {
  const T n = sqrt_t((v.x * v.x + v.y * v.y) + v.z * v.z);
  return V3<T>{v.x / n, v.y / n, v.z / n};
}
// Camera.inv_scale_translation_matrix (geometry.py:472-511)
template <typename T>
__device__ __forceinline__ void inv_scale_translation(const T* m, T* out) {
  T r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = T(1.0f) / m[5 * i];
  mat4_zero(out);
#pragma unroll
  for (int i = 0; i < 4; ++i) out[5 * i] = r[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) out[4 * i + 3] = (-(r[i] * m[4 * i + 3])) * r[3];
}

// All 8 matrices of one camera from its 16-parameter row (+ the viewport matrix in light mode), in the field order
// of `Camera`: view, projection, viewport, world_to_clip, world_to_eye_norm, world_to_screen, view_inv, screen_to_world.
template <typename T>
__device__ __forceinline__ void camera_matrices(int mode, const T* p, const T* v, T (*out)[16]) {
  V3<T> eye, centre{p[3], p[4], p[5]}, up{p[6], p[7], p[8]};
  T* proj = out[1];
  T* vp = out[2];
  mat4_zero(proj);
  if (mode == JR_CAMERA_PERSPECTIVE) {
    eye = V3<T>{p[0], p[1], p[2]};
    const T kRad(0.017453292519943295f);
    const T tv = tan_t((p[9] * kRad) / T(2.0f)), th = tan_t((p[10] * kRad) / T(2.0f));
    const T f = T(1.0f) / tv, aspect = th / tv, zn = p[11], zf = p[12];
    proj[0] = f / aspect;
    proj[5] = f;
    proj[10] = (zf + zn) / (zn - zf);
    proj[11] = ((T(2.0f) * zf) * zn) / (zn - zf);
    proj[14] = T(-1.0f);
    mat4_zero(vp);
    const T w = p[13], h = p[14], d = p[15];
    vp[0] = w / T(2.0f); vp[3] = T(0.0f) + w / T(2.0f);
    vp[5] = h / T(2.0f); vp[7] = T(0.0f) + h / T(2.0f);
    vp[10] = d / T(2.0f); vp[11] = d / T(2.0f);
    vp[15] = T(1.0f);
  } else {
    // eye = centre + light_direction * distance (shadow.py:84-88); here p[0..2] is the centre
    centre = V3<T>{p[0], p[1], p[2]};
    eye = V3<T>{centre.x + p[3] * p[9], centre.y + p[4] * p[9], centre.z + p[5] * p[9]};
    const T l = p[10], r = p[11], bo = p[12], t = p[13], n = p[14], fa = p[15];
    proj[0] = T(2.0f) / (r - l);
    proj[5] = T(2.0f) / (t - bo);
    proj[10] = T(-2.0f) / (fa - n);
    proj[15] = T(1.0f);
    proj[3] = -(r + l) / (r - l);
    proj[7] = -(t + bo) / (t - bo);
    proj[11] = -(fa + n) / (fa - n);
#pragma unroll
    for (int i = 0; i < 16; ++i) vp[i] = v[i];
  }
  // lookAt (geometry.py:536-575) and its analytic inverse (:577-636)
  const V3<T> fwd = normalise_t(V3<T>{centre.x - eye.x, centre.y - eye.y, centre.z - eye.z});
  const V3<T> upn = normalise_t(up);
  const V3<T> side = normalise_t(cross_t(fwd, upn));
  const V3<T> up2 = cross_t(side, fwd);
  const T R[9] = {side.x, side.y, side.z, up2.x, up2.y, up2.z, -fwd.x, -fwd.y, -fwd.z};
  T* view = out[0];
  T* view_inv = out[6];
  mat4_zero(view); mat4_zero(view_inv);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) { view[4 * i + j] = R[3 * i + j]; view_inv[4 * j + i] = R[3 * i + j]; }
    view[4 * i + 3] = -((R[3 * i] * eye.x + R[3 * i + 1] * eye.y) + R[3 * i + 2] * eye.z);
  }
  view_inv[3] = eye.x; view_inv[7] = eye.y; view_inv[11] = eye.z;
  view[15] = T(1.0f); view_inv[15] = T(1.0f);

  // projection inverse: perspective (columns / rows 2 and 3 swapped around the scale-translation
  // inverse, geometry.py:686-718) when projection[3][3] is ~0, orthographic otherwise
  T proj_inv[16], tmp[16], sh[16];
  if (fabsf(val(proj[15])) <= 1e-8f) {
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
  T vp_inv[16];
  inv_scale_translation(vp, vp_inv);

  mat4_mul(proj, view, out[3]);          // world_to_clip
  mat4_mul(vp, proj, tmp);
  mat4_mul(tmp, view, out[5]);           // world_to_screen
  mat4_mul(view_inv, proj_inv, tmp);
  mat4_mul(tmp, vp_inv, out[7]);         // screen_to_world
#pragma unroll
  for (int i = 0; i < 16; ++i) out[4][i] = view_inv[4 * (i & 3) + (i >> 2)];   // world_to_eye_norm = view_inv^T
}

__global__ void __launch_bounds__(128) k_camera(const __grid_constant__ JrCameraArgs a) {
  const int b = blockIdx.x * 128 + threadIdx.x;
  if (b >= a.B) return;
  const float* p = a.params.ptr + (long long)b * a.params.batch_stride;
  const float* v = a.viewport.ptr ? a.viewport.ptr + (long long)b * a.viewport.batch_stride : nullptr;
  float m[8][16];
  camera_matrices<float>(a.mode, p, v, m);
  const long long plane = (long long)a.B * 16;
  float* o = a.out + (long long)b * 16;
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int i = 0; i < 16; ++i) o[k * plane + i] = m[k][i];
}

// Reverse mode of k_camera: d_params[j] = sum over the 8 x 16 outputs of d_out * d out / d params[j],
// the derivative taken by running the SAME formulas on dual numbers, once per parameter (16 x ~700 flops per camera;
// one thread per (camera, parameter)).  Light mode: the viewport matrix is an input too (d_viewport[j] alike).
__global__ void __launch_bounds__(128) k_camera_vjp(const __grid_constant__ JrCameraArgs a, const float* __restrict__ d_out,
                                                    float* __restrict__ d_params, float* __restrict__ d_viewport) {
  const int n_in = (a.mode == JR_CAMERA_LIGHT && d_viewport) ? 32 : 16;
  const long long idx = (long long)blockIdx.x * 128 + threadIdx.x;
  if (idx >= (long long)a.B * n_in) return;
  const int b = (int)(idx / n_in), j = (int)(idx - (long long)b * n_in);
  const float* p = a.params.ptr + (long long)b * a.params.batch_stride;
  const float* v = a.viewport.ptr ? a.viewport.ptr + (long long)b * a.viewport.batch_stride : nullptr;
  Dual dp[16], dv[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    dp[i] = Dual(p[i], (j == i) ? 1.f : 0.f);
    dv[i] = Dual(v ? v[i] : 0.f, (j == 16 + i) ? 1.f : 0.f);
  }
  Dual m[8][16];
  camera_matrices<Dual>(a.mode, dp, dv, m);
  const long long plane = (long long)a.B * 16;
  const float* g = d_out + (long long)b * 16;
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += g[k * plane + i] * m[k][i].d;
  if (j < 16) d_params[(long long)b * 16 + j] = acc;
  else d_viewport[(long long)b * 16 + (j - 16)] = acc;
}

}  // namespace jr

extern "C" int jr_camera_build(const JrCameraArgs* a, jr_stream_t stream) {
  if (!a || !a->params.ptr || !a->out) return JR_ERR_NULL;
  if (a->B <= 0) return JR_ERR_DIMS;
  if (a->mode != JR_CAMERA_PERSPECTIVE && a->mode != JR_CAMERA_LIGHT) return JR_ERR_UNSUPPORTED;
  if (a->mode == JR_CAMERA_LIGHT && !a->viewport.ptr) return JR_ERR_NULL;
  jr::mark((cudaStream_t)stream, nullptr);
  jr::k_camera<<<(a->B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*a);
  jr::g_launches++;
  jr::mark((cudaStream_t)stream, "k_camera");
  return cudaGetLastError() == cudaSuccess ? JR_OK : JR_ERR_CUDA;
}

extern "C" int jr_camera_vjp(const JrCameraArgs* a, const float* d_out, float* d_params, float* d_viewport,
                             jr_stream_t stream) {
  if (!a || !a->params.ptr || !d_out || !d_params) return JR_ERR_NULL;
  if (a->B <= 0) return JR_ERR_DIMS;
  if (a->mode != JR_CAMERA_PERSPECTIVE && a->mode != JR_CAMERA_LIGHT) return JR_ERR_UNSUPPORTED;
  if (a->mode == JR_CAMERA_LIGHT && !a->viewport.ptr) return JR_ERR_NULL;
  const int n_in = (a->mode == JR_CAMERA_LIGHT && d_viewport) ? 32 : 16;
  const long long n = (long long)a->B * n_in;
  jr::mark((cudaStream_t)stream, nullptr);
  jr::k_camera_vjp<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*a, d_out, d_params, d_viewport);
  jr::g_launches++;
  jr::mark((cudaStream_t)stream, "k_camera_vjp");
  return cudaGetLastError() == cudaSuccess ? JR_OK : JR_ERR_CUDA;
}
