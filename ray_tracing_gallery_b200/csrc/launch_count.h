// launch_count.h — host-side tally of this library's own kernel launches (rt_kernel_launches()).
#pragma once
#include <atomic>
#include <cstdint>

namespace b200rt {
extern std::atomic<uint64_t> g_kernel_launches;
inline void note_launch(uint64_t n = 1) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace b200rt
