// micro-benchmark: issue throughput of FADD / FMUL / FFMA vs their packed FADD2 / FMUL2 / FFMA2 forms on sm_100a
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o f32x2 f32x2.cu && ./f32x2
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float s) {
  float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  float2 b0 = {a0, a1}, b1 = {a2, a3}, b2 = {a4, a5}, b3 = {a6, a7};
  const float2 s2 = {s, s};
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) { a0 += s; a1 += s; a2 += s; a3 += s; a4 += s; a5 += s; a6 += s; a7 += s; }
    if (MODE == 1) { b0 = __fadd2_rn(b0, s2); b1 = __fadd2_rn(b1, s2); b2 = __fadd2_rn(b2, s2); b3 = __fadd2_rn(b3, s2); }
    if (MODE == 2) { a0 *= s; a1 *= s; a2 *= s; a3 *= s; a4 *= s; a5 *= s; a6 *= s; a7 *= s; }
    if (MODE == 3) { b0 = __fmul2_rn(b0, s2); b1 = __fmul2_rn(b1, s2); b2 = __fmul2_rn(b2, s2); b3 = __fmul2_rn(b3, s2); }
    if (MODE == 4) { a0 = fmaf(a0, s, s); a1 = fmaf(a1, s, s); a2 = fmaf(a2, s, s); a3 = fmaf(a3, s, s); a4 = fmaf(a4, s, s); a5 = fmaf(a5, s, s); a6 = fmaf(a6, s, s); a7 = fmaf(a7, s, s); }
    if (MODE == 5) { b0 = __ffma2_rn(b0, s2, s2); b1 = __ffma2_rn(b1, s2, s2); b2 = __ffma2_rn(b2, s2, s2); b3 = __ffma2_rn(b3, s2, s2); }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + b0.x + b0.y + b1.x + b1.y + b2.x + b2.y + b3.x + b3.y;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 200000;
  const char* names[6] = {"FADD x8", "FADD2 x4", "FMUL x8", "FMUL2 x4", "FFMA x8", "FFMA2 x4"};
  for (int m = 0; m < 6; ++m) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      switch (m) {
        case 0: k<0><<<148 * 8, 256>>>(out, iters, 1.0001f); break;
        case 1: k<1><<<148 * 8, 256>>>(out, iters, 1.0001f); break;
        case 2: k<2><<<148 * 8, 256>>>(out, iters, 1.0001f); break;
        case 3: k<3><<<148 * 8, 256>>>(out, iters, 1.0001f); break;
        case 4: k<4><<<148 * 8, 256>>>(out, iters, 1.0001f); break;
        case 5: k<5><<<148 * 8, 256>>>(out, iters, 1.0001f); break;
      }
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 8.0 * iters * 148 * 8 * 256 * (m >= 4 ? 2 : 1);
    printf("%-9s %8.3f ms  %7.2f T%s/s (8 scalar results per iteration per thread)\n", names[m], ms, flops / ms / 1e9, m >= 4 ? "FLOP" : "op");
  }
  return 0;
}
