// launch_count.h -- counts the kernels this library launches (mq_launch_count(), bench.py's gpu_launches).
#pragma once
#include <atomic>
#include <stdint.h>
namespace mq {
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace mq
