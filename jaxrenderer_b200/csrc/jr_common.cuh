// jr_common.cuh -- host-side bits shared by the translation units.
#pragma once
#include <atomic>

namespace jr {
// kernels launched by this library since load (jr_launch_count()).
extern std::atomic<long long> g_launches;
}  // namespace jr
