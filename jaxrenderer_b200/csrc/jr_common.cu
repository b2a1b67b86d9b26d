#include "jr_common.cuh"
namespace jr {
std::atomic<long long> g_launches{0};
}
