// jr_common.cuh -- host-side bits shared by the translation units.
#pragma once
#include <atomic>
#include <cuda_runtime.h>

namespace jr {
// kernels launched by this library since load (jr_launch_count()).
extern std::atomic<long long> g_launches;

// Per-kernel timing (measurement aid, jr_debug_kernel_timing / jr_debug_kernel_times in jr_b200.h): while enabled,
// every entry point records a CUDA event on its stream when it starts (name == nullptr) and after each launch (the
// kernel's name); a kernel's time is the distance to the event before it on the same call.  Off: one relaxed load.
extern std::atomic<int> g_timing;
void timing_mark(cudaStream_t stream, const char* name);
inline void mark(cudaStream_t stream, const char* name) {
  if (g_timing.load(std::memory_order_relaxed)) timing_mark(stream, name);
}
}  // namespace jr
