// Multi-GPU assembly of intertable shards over NVLink peer memory (north_star: "all-gather ... only to
// assemble table shards").  Every rank owns the table rows of a contiguous block of targets; the full table lives in a
// buffer that is mapped into every process of the node (torch symmetric memory / CUDA IPC), at the same layout on every
// GPU.  A rank writes its block into its own copy (g2p_knn_block) and this kernel stores the same bytes into the peers'
// copies - plain 16-byte stores through the NVLink mapping, one launch for all peers and all table rows, running on a
// second stream while the k-NN of the rank's next target block computes.  No staging buffer, no library collective: the
// only other exchange is the barrier that tells every rank the table is complete.
#include <cstdlib>

#include "g2p_common.cuh"

namespace g2p {

constexpr int kMaxPeers = 15;
struct PeerList {
  char* base[kMaxPeers];
};

// rows of `row_bytes` bytes; the segment [col_off, col_off + seg_bytes) of every row goes to all peers.
// grid.y = row, grid.z = peer, grid.x strides over the segment.  VEC = bytes per access (16 or 4).
template <int VEC>
__global__ void __launch_bounds__(256) peer_push_kernel(const char* __restrict__ local, PeerList peers, int64_t row_bytes,
                                                        int64_t col_off, int64_t seg_bytes) {
  const int64_t at = (int64_t)blockIdx.y * row_bytes + col_off;
  const char* src = local + at;
  char* dst = peers.base[blockIdx.z] + at;
  const int64_t nvec = seg_bytes / VEC;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (VEC == 16) {
    // four independent 16-byte loads in flight per thread: the few CTAs this kernel is given (it shares the SMs with
    // the k-NN of the next target block) must cover the HBM latency on their own
    for (; i + 3 * step < nvec; i += 4 * step) {
      const int4 v0 = __ldg(reinterpret_cast<const int4*>(src) + i);
      const int4 v1 = __ldg(reinterpret_cast<const int4*>(src) + i + step);
      const int4 v2 = __ldg(reinterpret_cast<const int4*>(src) + i + 2 * step);
      const int4 v3 = __ldg(reinterpret_cast<const int4*>(src) + i + 3 * step);
      reinterpret_cast<int4*>(dst)[i] = v0;
      reinterpret_cast<int4*>(dst)[i + step] = v1;
      reinterpret_cast<int4*>(dst)[i + 2 * step] = v2;
      reinterpret_cast<int4*>(dst)[i + 3 * step] = v3;
    }
    for (; i < nvec; i += step) reinterpret_cast<int4*>(dst)[i] = __ldg(reinterpret_cast<const int4*>(src) + i);
  } else {
    for (; i < nvec; i += step) reinterpret_cast<int32_t*>(dst)[i] = __ldg(reinterpret_cast<const int32_t*>(src) + i);
  }
}

// the same rows stored ONCE to the multicast address of the buffer: the NVSwitch replicates every 16-byte store into all
// GPUs of the group (NVLS; this GPU's own copy included), so a rank sends its block once instead of once per peer
__global__ void __launch_bounds__(256) multicast_push_kernel(const char* __restrict__ local, char* __restrict__ mc,
                                                             int64_t row_bytes, int64_t col_off, int64_t seg_bytes) {
  const int64_t at = (int64_t)blockIdx.y * row_bytes + col_off;
  const int4* src = reinterpret_cast<const int4*>(local + at);
  char* dst = mc + at;
  const int64_t nvec = seg_bytes / 16;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  auto mc_store = [&](int64_t i, const int4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + 16 * i), "f"(__int_as_float(v.x)),
                 "f"(__int_as_float(v.y)), "f"(__int_as_float(v.z)), "f"(__int_as_float(v.w))
                 : "memory");
  };
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * step < nvec; i += 4 * step) {
    const int4 v0 = __ldg(src + i), v1 = __ldg(src + i + step), v2 = __ldg(src + i + 2 * step), v3 = __ldg(src + i + 3 * step);
    mc_store(i, v0);
    mc_store(i + step, v1);
    mc_store(i + 2 * step, v2);
    mc_store(i + 3 * step, v3);
  }
  for (; i < nvec; i += step) mc_store(i, __ldg(src + i));
}

}  // namespace g2p

using namespace g2p;

extern "C" int g2p_multicast_push(const void* local_base, void* multicast_base, int64_t row_bytes, int64_t col_off,
                                  int64_t seg_bytes, int n_rows, void* stream) {
  if (!local_base || !multicast_base) return set_error(G2P_ERR_INVALID, "g2p_multicast_push: null pointer");
  if (row_bytes < 0 || col_off < 0 || seg_bytes < 0 || col_off + seg_bytes > row_bytes || n_rows < 0)
    return set_error(G2P_ERR_INVALID, "g2p_multicast_push: bad sizes");
  if ((((uintptr_t)local_base | (uintptr_t)multicast_base | (uintptr_t)row_bytes | (uintptr_t)col_off | (uintptr_t)seg_bytes) & 15) != 0)
    return set_error(G2P_ERR_UNSUPPORTED, "g2p_multicast_push: 16-byte aligned buffers, rows and segments only (use g2p_peer_push)");
  if (n_rows == 0 || seg_bytes == 0) return G2P_OK;
  static const int env_ctas = [] {
    const char* e = getenv("G2P_PUSH_CTAS");
    return e ? atoi(e) : 0;
  }();
  const int want_ctas = env_ctas > 0 ? env_ctas : 32;
  int gx = (int)std::min<int64_t>(ceil_div(seg_bytes / 16, 256 * 4), std::max(1, want_ctas / std::max(1, n_rows)));
  if (gx < 1) gx = 1;
  multicast_push_kernel<<<dim3((unsigned)gx, (unsigned)n_rows), 256, 0, as_stream(stream)>>>(
      static_cast<const char*>(local_base), static_cast<char*>(multicast_base), row_bytes, col_off, seg_bytes);
  G2P_LAUNCHED();
  return G2P_OK;
}

extern "C" int g2p_peer_push(const void* local_base, void* const* peer_bases, int n_peers, int64_t row_bytes,
                             int64_t col_off, int64_t seg_bytes, int n_rows, void* stream) {
  if (!local_base || (n_peers > 0 && !peer_bases)) return set_error(G2P_ERR_INVALID, "g2p_peer_push: null pointer");
  if (n_peers < 0 || n_peers > kMaxPeers) return set_error(G2P_ERR_INVALID, "g2p_peer_push: 0..%d peers", kMaxPeers);
  if (row_bytes < 0 || col_off < 0 || seg_bytes < 0 || col_off + seg_bytes > row_bytes || n_rows < 0)
    return set_error(G2P_ERR_INVALID, "g2p_peer_push: bad sizes");
  if ((row_bytes | col_off | seg_bytes) & 3) return set_error(G2P_ERR_INVALID, "g2p_peer_push: sizes must be multiples of 4 bytes");
  if (n_peers == 0 || n_rows == 0 || seg_bytes == 0) return G2P_OK;
  // G2P_PUSH_MODE=ce: the copy engines instead of SMs (one strided DMA per peer); the default is the store kernel
  static const bool use_ce = [] {
    const char* e = getenv("G2P_PUSH_MODE");
    return e && e[0] == 'c';
  }();
  if (use_ce) {
    for (int p = 0; p < n_peers; ++p)
      G2P_CUDA(cudaMemcpy2DAsync(static_cast<char*>(peer_bases[p]) + col_off, (size_t)row_bytes,
                                 static_cast<const char*>(local_base) + col_off, (size_t)row_bytes, (size_t)seg_bytes,
                                 (size_t)n_rows, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return G2P_OK;
  }
  PeerList pl{};
  bool v16 = (((uintptr_t)local_base | (uintptr_t)row_bytes | (uintptr_t)col_off | (uintptr_t)seg_bytes) & 15) == 0;
  for (int p = 0; p < n_peers; ++p) {
    if (!peer_bases[p]) return set_error(G2P_ERR_INVALID, "g2p_peer_push: peer %d has no mapping", p);
    pl.base[p] = static_cast<char*>(peer_bases[p]);
    v16 = v16 && (((uintptr_t)peer_bases[p]) & 15) == 0;
  }
  const int vec = v16 ? 16 : 4;
  const int64_t nvec = seg_bytes / vec;
  // a small grid over all rows and peers (G2P_PUSH_CTAS; default 32 CTAs, 56 for >= 4 peers): enough to fill the
  // NVLink ports (measured 800 GB/s per GPU at 8 GPUs), few enough to leave the SMs to the k-NN kernel this push
  // overlaps with (with 148+ CTAs the overlap is lost: 3.18 against 2.83 ms per 2-GPU table assembly)
  static const int env_ctas = [] {
    const char* e = getenv("G2P_PUSH_CTAS");
    return e ? atoi(e) : 0;
  }();
  const int want_ctas = env_ctas > 0 ? env_ctas : (n_peers >= 4 ? 56 : 32);
  int gx = (int)std::min<int64_t>(ceil_div(nvec, 256 * 4), std::max(1, want_ctas / std::max(1, n_rows * n_peers)));
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)n_rows, (unsigned)n_peers);
  if (v16)
    peer_push_kernel<16><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const char*>(local_base), pl, row_bytes, col_off, seg_bytes);
  else
    peer_push_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const char*>(local_base), pl, row_bytes, col_off, seg_bytes);
  G2P_LAUNCHED();
  return G2P_OK;
}
