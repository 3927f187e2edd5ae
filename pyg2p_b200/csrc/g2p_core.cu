// Error handling, device queries and launch accounting for libg2pgpu.so.
#include <cstdarg>
#include <cstdio>
#include <cstring>

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

// Host-side staging helper of the ingest path (no device work): out[c] = src[touched[c]] for the nc source points a
// table reads, straight from the pageable array the GRIB decoder returned into a (pinned) compact buffer; optionally
// the mask bytes of those points (out_mask) and / or the marker NaN folded into the values of masked points.
int g2p_host_pick(const double* src, const uint8_t* mask, const int32_t* touched, int64_t nc, double* out,
                  uint8_t* out_mask, int fold_marker) {
  if (!src || !touched || !out || nc < 0) return g2p::set_error(G2P_ERR_INVALID, "g2p_host_pick: null pointer");
  const double marker = [] {
    const unsigned long long u = G2P_MASK_MARKER;
    double d;
    memcpy(&d, &u, sizeof d);
    return d;
  }();
  if (!mask) {
    for (int64_t c = 0; c < nc; ++c) out[c] = src[touched[c]];
    if (out_mask) memset(out_mask, 0, (size_t)nc);
    return G2P_OK;
  }
  for (int64_t c = 0; c < nc; ++c) {
    const int32_t j = touched[c];
    const uint8_t mk = mask[j] ? 1 : 0;
    out[c] = (mk && fold_marker) ? marker : src[j];
    if (out_mask) out_mask[c] = mk;
  }
  return G2P_OK;
}

}  // extern "C"
