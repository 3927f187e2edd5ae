// Error handling, device queries and launch accounting for libg2pgpu.so.
#include <cstdarg>
#include <cstdio>

#include "g2p_common.cuh"

namespace g2p {

static thread_local char t_err[512] = "";
std::atomic<int64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  int code = (e == cudaErrorMemoryAllocation) ? G2P_ERR_NOMEM
             : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? G2P_ERR_NO_DEVICE
                                                                            : G2P_ERR_CUDA;
  return set_error(code, "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace g2p

extern "C" {

int g2p_version(void) { return G2P_VERSION; }

const char* g2p_last_error(void) { return g2p::t_err; }

int g2p_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return g2p::cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
  return n;
}

int g2p_set_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return g2p::cuda_fail(e, "cudaSetDevice", __FILE__, __LINE__);
  return G2P_OK;
}

int64_t g2p_launch_count(void) { return g2p::g_launches.load(); }

}  // extern "C"
