// Count of kernel launches issued by this library (taco_launch_count): bench.py reports it as `gpu_launches`.
#pragma once
#include <atomic>
namespace taco { extern std::atomic<unsigned long long> g_launches; }
#define TACO_LAUNCHED() (void)::taco::g_launches.fetch_add(1ull, std::memory_order_relaxed)
