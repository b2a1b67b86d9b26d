#include "jr_common.cuh"

#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/jr_b200.h"

namespace jr {
std::atomic<long long> g_launches{0};
std::atomic<int> g_timing{0};

namespace {
struct Mark {
  cudaEvent_t ev;
  const char* name;  // nullptr: start of an entry point
};
std::mutex g_marks_mutex;
std::vector<Mark> g_marks;

void clear_marks() {
  for (Mark& m : g_marks) cudaEventDestroy(m.ev);
  g_marks.clear();
}
}  // namespace

void timing_mark(cudaStream_t stream, const char* name) {
  std::lock_guard<std::mutex> lock(g_marks_mutex);
  if (g_marks.size() >= (size_t)JR_KERNEL_TIMES_MAX) return;  // bounded: a forgotten switch cannot eat the host
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  if (cudaEventRecord(ev, stream) != cudaSuccess) { cudaEventDestroy(ev); return; }
  g_marks.push_back(Mark{ev, name});
}
}  // namespace jr

extern "C" {

int jr_debug_kernel_timing(int enable) {
  std::lock_guard<std::mutex> lock(jr::g_marks_mutex);
  jr::clear_marks();
  jr::g_timing.store(enable ? 1 : 0);
  return JR_OK;
}

int jr_debug_kernel_times(JrKernelTime* out, int capacity) {
  if (capacity > 0 && !out) return JR_ERR_NULL;
  std::lock_guard<std::mutex> lock(jr::g_marks_mutex);
  int n = 0, call = -1;
  for (size_t i = 0; i < jr::g_marks.size(); ++i) {
    const jr::Mark& m = jr::g_marks[i];
    if (!m.name) { ++call; continue; }
    if (i == 0) continue;  // a launch with no start mark before it (switched on mid-call): no interval to report
    if (n < capacity) {
      if (cudaEventSynchronize(m.ev) != cudaSuccess) return JR_ERR_CUDA;
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, jr::g_marks[i - 1].ev, m.ev) != cudaSuccess) return JR_ERR_CUDA;
      std::strncpy(out[n].name, m.name, sizeof(out[n].name) - 1);
      out[n].name[sizeof(out[n].name) - 1] = '\0';
      out[n].ms = ms;
      out[n].call = call < 0 ? 0 : call;
    }
    ++n;
  }
  jr::clear_marks();
  return n;
}

}  // extern "C"
