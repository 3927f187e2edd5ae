// Shared helpers for libg2pgpu.so (sm_100a).  See include/g2p.h for the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <functional>

#include "../../include/g2p.h"

namespace g2p {

int set_error(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
extern std::atomic<int64_t> g_launches;
int num_sms();
// f(lo, hi) over [0, n) in a few slices on persistent helper threads (host-side staging loops)
void host_parallel_for(int64_t n, int64_t min_slice, int threads, const std::function<void(int64_t, int64_t)>& f);

#define G2P_CUDA(expr)                                                     \
  do {                                                                     \
    cudaError_t _e = (expr);                                               \
    if (_e != cudaSuccess) return g2p::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

// call after every kernel launch: counts it and surfaces launch-configuration errors
#define G2P_LAUNCHED()                                                     \
  do {                                                                     \
    g2p::g_launches.fetch_add(1, std::memory_order_relaxed);               \
    cudaError_t _e = cudaGetLastError();                                   \
    if (_e != cudaSuccess) return g2p::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace g2p
