// jr_device.cuh -- exact-order fp32 device math shared by the visibility,
// shading and backward kernels.
//
// This translation unit is compiled with -fmad=false: every `*`, `+`, `-` below
// is ONE rounded fp32 operation and `/`, sqrtf are IEEE round-to-nearest, so the
// discrete outcomes (edge inclusion, depth order, texel choice) are bit-equal
// to the oracle's scalar-order restatement of the reference
// (renderer/pipeline.py:76-113, :163-279; geometry.py:71-110, :284-315).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace jr {

struct Vec3 { float x, y, z; };

__device__ __forceinline__ float dot3(float a0, float a1, float a2, float b0, float b1, float b2) {
  return (a0 * b0 + a1 * b1) + a2 * b2;
}
// IEEE a / b, bit for bit.  A ZERO dividend sends the compiler's division through its out-of-line slow path (FCHK
// flags it: ~35 extra instructions per warp that holds one -- axis-aligned normals, the centre row / column of the
// canvas); +-0 / b is +-0 for every finite non-zero b, so that case is answered directly and the division sees 1 / b.
// (The substituted dividend goes through an empty asm: otherwise the compiler sees that the quotient is only used
// when the dividend was not replaced, divides the original value after all, and the zero reaches FCHK again.)
__device__ __forceinline__ float opaque(float x) {
  asm volatile("" : "+f"(x));
  return x;
}
__device__ __forceinline__ float fdiv_z(float a, float b) {
  const bool z = a == 0.f && fabsf(b) > 0.f && fabsf(b) <= 3.4028234664e38f;
  const float q = opaque(z ? 1.f : a) / b;
  return z ? __uint_as_float((__float_as_uint(a) ^ __float_as_uint(b)) & 0x80000000u) : q;
}
__device__ __forceinline__ Vec3 normalise3(Vec3 v) {
  float n = sqrtf(dot3(v.x, v.y, v.z, v.x, v.y, v.z));
  // v / n with the zero components answered directly (see fdiv_z; n >= 0 here, so +-0 / n keeps its sign)
  const bool ok = n > 0.f && n <= 3.4028234664e38f;
  const bool zx = ok && v.x == 0.f, zy = ok && v.y == 0.f, zz = ok && v.z == 0.f;
  const float qx = opaque(zx ? 1.f : v.x) / n, qy = opaque(zy ? 1.f : v.y) / n, qz = opaque(zz ? 1.f : v.z) / n;
  return Vec3{zx ? v.x : qx, zy ? v.y : qy, zz ? v.z : qz};
}

// v / |v| for quantities NOTHING DISCRETE depends on (the interpolated shading normal, the reflection vector: they feed
// the continuous lighting terms only; BASELINE's contract for colours is 1e-5): reciprocal square root + one Newton
// step (< 1 ulp) and three products instead of an IEEE square root and three IEEE quotients -- 12 instead of ~55
// instructions.  Everything that can decide a texel, a shadow test, a discarded fragment or a triangle keeps the
// exact normalise3 above.  A zero vector gives NaN, as 0 / 0 does.
__device__ __forceinline__ Vec3 normalise3_fast(Vec3 v) {
  const float d = fmaf(v.x, v.x, fmaf(v.y, v.y, v.z * v.z));
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  r = r * fmaf(-0.5f * d, r * r, 1.5f);
  return Vec3{v.x * r, v.y * r, v.z * r};
}

// to_homogeneous(p) @ M.T  (Camera.to_clip, geometry.py:420-438)
__device__ __forceinline__ void to_clip(const float* __restrict__ M, float x, float y, float z,
                                        float out[4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
    out[r] = ((x * M[4 * r + 0] + y * M[4 * r + 1]) + z * M[4 * r + 2]) + M[4 * r + 3];
}

// Camera.apply_vec (geometry.py:352-389): normalise, rotate (w = 0), normalise.
__device__ __forceinline__ Vec3 apply_vec(const float* __restrict__ M, Vec3 v) {
  Vec3 n = normalise3(v);
  Vec3 t;
  t.x = (n.x * M[0] + n.y * M[1]) + n.z * M[2];
  t.y = (n.x * M[4] + n.y * M[5]) + n.z * M[6];
  t.z = (n.x * M[8] + n.y * M[9]) + n.z * M[10];
  return normalise3(t);
}

// jnp.linalg.det closed form, jax `_det_3x3` term order (pipeline.py:91-94).
__device__ __forceinline__ float det3(const float a[9]) {
  return a[0] * a[4] * a[8] + a[1] * a[5] * a[6] + a[2] * a[3] * a[7] - a[2] * a[4] * a[6] -
         a[0] * a[5] * a[7] - a[1] * a[3] * a[8];
}

__device__ __forceinline__ void swapf(float& a, float& b) { float t = a; a = b; b = t; }

// jnp.linalg.inv of a 3x3 (pipeline.py:105): LU with partial pivoting in LAPACK
// sgetrf2 order (column scaled by the reciprocal pivot) and strsm-ordered
// substitutions against the permuted identity.  Mirrors oracle.lu_inverse3.
__device__ __forceinline__ void lu_inverse3(const float A[9], float inv[9]) {
  float r0[6] = {A[0], A[1], A[2], 1.f, 0.f, 0.f};
  float r1[6] = {A[3], A[4], A[5], 0.f, 1.f, 0.f};
  float r2[6] = {A[6], A[7], A[8], 0.f, 0.f, 1.f};
  float a0 = fabsf(r0[0]), a1 = fabsf(r1[0]), a2 = fabsf(r2[0]);
  bool p1 = a1 > a0;
  float best = p1 ? a1 : a0;
  bool p2 = a2 > best;
  p1 = p1 && !p2;
  if (p1) {
#pragma unroll
    for (int c = 0; c < 6; ++c) swapf(r0[c], r1[c]);
  }
  if (p2) {
#pragma unroll
    for (int c = 0; c < 6; ++c) swapf(r0[c], r2[c]);
  }
  float r00 = 1.0f / r0[0];
  float l10 = r1[0] * r00, l20 = r2[0] * r00;
  float u00 = r0[0], u01 = r0[1], u02 = r0[2];
  float a11 = r1[1] - l10 * u01, a12 = r1[2] - l10 * u02;
  float a21 = r2[1] - l20 * u01, a22 = r2[2] - l20 * u02;
  if (fabsf(a21) > fabsf(a11)) {
    swapf(l10, l20); swapf(a11, a21); swapf(a12, a22);
#pragma unroll
    for (int c = 3; c < 6; ++c) swapf(r1[c], r2[c]);
  }
  float u11 = a11, u12 = a12;
  float l21 = a21 * (1.0f / u11);
  float u22 = a22 - l21 * u12;
#ifdef JR_LU_ROLLED
#pragma unroll 1
#else
#pragma unroll
#endif
  for (int j = 0; j < 3; ++j) {
    float y0 = r0[3 + j];
    float y1 = r1[3 + j] - y0 * l10;
    float y2 = (r2[3 + j] - y0 * l20) - y1 * l21;
    float x2 = y2 / u22;
    float t1 = y1 - x2 * u12;
    float t0 = y0 - x2 * u02;
    float x1 = t1 / u11;
    t0 = t0 - x1 * u01;
    float x0 = t0 / u00;
    inv[0 + j] = x0; inv[3 + j] = x1; inv[6 + j] = x2;
  }
}

// Monotone map float -> uint32 (-0 canonicalised to +0 first).
__device__ __forceinline__ uint32_t orderable(float z) {
  const uint32_t f = __float_as_uint(z + 0.0f);
  return f ^ ((uint32_t)((int32_t)f >> 31) | 0x80000000u);  // negative: flip all bits; else: set the sign bit
}
__device__ __forceinline__ float from_orderable(uint32_t u) {
  return __uint_as_float(u ^ (~(uint32_t)((int32_t)u >> 31) | 0x80000000u));
}

// Per-triangle raster record (PerPrimitive, pipeline.py:49-113).
struct TriSetup {
  float inv[9];  // matrix_inv, row-major [row][col]
  float zc[3];   // clip-space z of the three vertices
  float det;
};

// Build PerPrimitive from three clip-space vertices.  Returns det.
__device__ __forceinline__ float tri_matrix(const float c0[4], const float c1[4], const float c2[4],
                                            float M[9]) {
  M[0] = c0[0]; M[1] = c0[1]; M[2] = c0[3];
  M[3] = c1[0]; M[4] = c1[1]; M[5] = c1[3];
  M[6] = c2[0]; M[7] = c2[1]; M[8] = c2[3];
  return det3(M);
}

// Edge functions / clip_coef at NDC pixel (xn, yn) (pipeline.py:190).
__device__ __forceinline__ void clip_coef(const float inv[9], float xn, float yn, float c[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) c[k] = (xn * inv[k] + yn * inv[3 + k]) + inv[6 + k];
}

__device__ __forceinline__ float interp3(const float tc[3], float v0, float v1, float v2) {
  return (tc[0] * v0 + tc[1] * v1) + tc[2] * v2;
}

__device__ __forceinline__ int pymod(int a, int n) {
  int m = a % n;
  return m < 0 ? m + n : m;
}
// jnp arr[i]: negative wraps once, then clamp.
__device__ __forceinline__ int wrap_clamp(int i, int n) {
  if (i < 0) i += n;
  return i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
}


// Conservative bbox margin (pixels).  The fp32 edge functions can include a pixel centre that lies
// slightly OUTSIDE the exact triangle; the bbox must keep every such pixel.  The excess distance is
// bounded by ~ eps * W * (L/h)^2 pixels (evaluation error eps*W, amplified by the conditioning L/h of
// the 3x3 inverse and by 1/sin of the sharpest corner ~ L/h; L = longest edge, h = smallest height).
// Small, well-shaped triangles therefore get a margin of 1/64 px (with a 64x safety factor on the
// estimate), growing to the blanket 0.5 px for needles and for anything larger than 8 px (where the
// margin costs little).  About half of the candidate triangles of a Brax scene at 84x84 contain no
// sample at all and are dropped here, before the LU inverse (validated bit-for-bit against the
// brute-force oracle on millions of random small / needle triangles: tests/test_gpu_fuzz.py).
__device__ __forceinline__ float bbox_margin(float sx0, float sy0, float sx1, float sy1, float sx2, float sy2,
                                             float view_w) {
  const float ex = fmaxf(fmaxf(sx0, sx1), sx2) - fminf(fminf(sx0, sx1), sx2);
  const float ey = fmaxf(fmaxf(sy0, sy1), sy2) - fminf(fminf(sy0, sy1), sy2);
  if (!(ex < 8.f && ey < 8.f)) return 0.5f;  // also NaN
  const float ax = sx1 - sx0, ay = sy1 - sy0, bx = sx2 - sx0, by = sy2 - sy0, cx = sx2 - sx1, cy = sy2 - sy1;
  const float l2 = fmaxf(fmaxf(ax * ax + ay * ay, bx * bx + by * by), cx * cx + cy * cy);
  const float area2 = fabsf(ax * by - ay * bx);           // twice the area
  const float asp = l2 / fmaxf(area2, 1e-20f);            // L / h
  const float m = 3.8e-6f * view_w * asp * asp;           // 64 * eps * W * (L/h)^2
  return fminf(0.5f, fmaxf(m, 0.015625f));
}

}  // namespace jr
