// Error handling, device queries and launch accounting for libg2pgpu.so.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

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

// ---- a few persistent helper threads for host-side staging loops (g2p_host_pick).  One job at a time per process
// (callers serialise on the job mutex); the calling thread takes the first slice itself.
namespace {
struct HostPool {
  std::mutex job_mutex;                 // one parallel_for at a time
  std::mutex m;
  std::condition_variable cv_work, cv_done;
  std::vector<std::thread> workers;
  const std::function<void(int64_t, int64_t)>* fn = nullptr;
  int64_t n = 0, slice = 0;
  int n_slices = 0, next = 0, pending = 0;
  uint64_t generation = 0;
  bool stop = false;

  explicit HostPool(int helpers) {
    for (int i = 0; i < helpers; ++i) workers.emplace_back([this] { loop(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> g(m);
      stop = true;
    }
    cv_work.notify_all();
    for (auto& t : workers) t.join();
  }
  void run_slices(std::unique_lock<std::mutex>& lk) {
    while (next < n_slices) {
      const int s = next++;
      lk.unlock();
      const int64_t lo = (int64_t)s * slice, hi = std::min(n, lo + slice);
      (*fn)(lo, hi);
      lk.lock();
      if (--pending == 0) cv_done.notify_all();
    }
  }
  void loop() {
    std::unique_lock<std::mutex> lk(m);
    uint64_t seen = 0;
    for (;;) {
      cv_work.wait(lk, [&] { return stop || (generation != seen && next < n_slices); });
      if (stop) return;
      seen = generation;
      run_slices(lk);
    }
  }
  void parallel_for(int64_t total, int64_t step, const std::function<void(int64_t, int64_t)>& f) {
    std::lock_guard<std::mutex> job(job_mutex);
    std::unique_lock<std::mutex> lk(m);
    fn = &f;
    n = total;
    slice = step;
    n_slices = (int)((total + step - 1) / step);
    next = 0;
    pending = n_slices;
    ++generation;
    // one helper per extra slice: waking sleepers that find no slice left costs more than their help is worth (a
    // 7-helper pool woken as a whole: 0.45 ms per streamed message, 3 helpers: 0.29)
    for (int i = 1; i < n_slices && i <= (int)workers.size(); ++i) cv_work.notify_one();
    run_slices(lk);
    cv_done.wait(lk, [&] { return pending == 0; });
    fn = nullptr;
  }
};
}  // namespace

void host_parallel_for(int64_t n, int64_t min_slice, int threads, const std::function<void(int64_t, int64_t)>& f) {
  constexpr int kMaxThreads = 4;
  static const int dflt = [] {
    const char* e = getenv("G2P_HOST_PICK_THREADS");
    const int v = e ? atoi(e) : 1;
    return v < 1 ? 1 : (v > kMaxThreads ? kMaxThreads : v);
  }();
  if (threads <= 0) threads = dflt;
  if (threads > kMaxThreads) threads = kMaxThreads;
  if (n <= 0) return;
  if (threads == 1 || n < 2 * min_slice) {
    f(0, n);
    return;
  }
  static HostPool pool(kMaxThreads - 1);          // the helpers sleep until a job has slices for them
  const int64_t parts = std::min<int64_t>(threads, n / min_slice);
  pool.parallel_for(n, (n + parts - 1) / parts, f);
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
// The gather is bound by the latency of the host memory (276 589 picks out of a 52.8 MB field: 0.28 ms on one core), so
// large picks can be split over a few persistent helper threads: `threads` <= 0 = the default (G2P_HOST_PICK_THREADS,
// else 1, the caller alone: a blocking per-message call loses with helpers - 0.79 / 0.86 / 0.88 ms with 1 / 2 / 4 threads,
// waking sleepers costs more than they save), while a stream of messages, whose helpers stay warm, is fastest with 4
// (0.44 -> 0.23 ms per message).
int g2p_host_pick(const double* src, const uint8_t* mask, const int32_t* touched, int64_t nc, double* out,
                  uint8_t* out_mask, int fold_marker, int threads) {
  if (!src || !touched || !out || nc < 0) return g2p::set_error(G2P_ERR_INVALID, "g2p_host_pick: null pointer");
  const double marker = [] {
    const unsigned long long u = G2P_MASK_MARKER;
    double d;
    memcpy(&d, &u, sizeof d);
    return d;
  }();
  auto part = [=](int64_t lo, int64_t hi) {
    if (!mask) {
      for (int64_t c = lo; c < hi; ++c) out[c] = src[touched[c]];
      if (out_mask) memset(out_mask + lo, 0, (size_t)(hi - lo));
      return;
    }
    for (int64_t c = lo; c < hi; ++c) {
      const int32_t j = touched[c];
      const uint8_t mk = mask[j] ? 1 : 0;
      out[c] = (mk && fold_marker) ? marker : src[j];
      if (out_mask) out_mask[c] = mk;
    }
  };
  g2p::host_parallel_for(nc, 32768, threads, part);
  return G2P_OK;
}
}  // extern "C"
