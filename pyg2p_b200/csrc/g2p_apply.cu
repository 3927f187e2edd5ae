// K6 (table record pack/unpack) and K7 (batched intertable apply) for sm_100a.
//
// K7 replaces Interpolator._interpolate_scipy_append_mv (reference
// src/pyg2p/main/interpolation/__init__.py:142-162), which gathers / einsum's ONE message per
// call on the CPU.  Here one launch applies the table to `b` messages:
//   - a thread owns two adjacent targets and keeps their k indexes + k weights in registers,
//     so every index/weight is loaded once per message chunk, not once per message;
//   - the message loop is unrolled x4: 4*2*k independent gathers are in flight per thread;
//   - outputs are written with 128-bit streaming stores (st.global.cs), gathers go through
//     the read-only path so that neighbouring targets share L1/L2 sectors;
//   - the grid is persistent (SMs x resident CTAs) and walks (target tile, message chunk)
//     tiles, target tile fastest, so concurrent CTAs stream one contiguous slab per message.
// Arithmetic is the reference's load-path order ((w0*v0 + w1*v1) + w2*v2) + w3*v3 with separate
// IEEE multiply and add (no FMA contraction) -> bit-exact.
#include "g2p_common.cuh"

namespace g2p {

constexpr int kApplyThreads = 256;
constexpr int kTargetsPerThread = 2;
constexpr int kMsgUnroll = 4;

struct ApplyArgs {
  const int32_t* idx;
  const double* coef;
  const double* src;
  const uint8_t* src_mask;
  double* out;
  uint8_t* out_mask;
  int64_t m, n, src_stride, out_stride;
  int64_t sbits_stride, obits_stride;  // bit-packed mask rows: bytes per source row, uint32 words per output row
  int b, k, pairwise;
  int msgs_per_chunk, n_target_tiles;
  int64_t n_tiles;
  double mv_out;
  int vec_store;  // out rows are 16-byte aligned for even targets
};

template <int K>
__device__ __forceinline__ double weighted(const double (&w)[K], const double (&v)[K], int pairwise) {
  if (K == 1) return v[0];
  if (K == 4 && pairwise)   // G2P_SUM_PAIRWISE
    return __dadd_rn(__dadd_rn(__dmul_rn(w[0], v[0]), __dmul_rn(w[K > 2 ? 2 : 0], v[K > 2 ? 2 : 0])),
                     __dadd_rn(__dmul_rn(w[1], v[1]), __dmul_rn(w[K > 3 ? 3 : 0], v[K > 3 ? 3 : 0])));
  if (K > 4 && pairwise) {
    // the order of numpy's buffered einsum for any k (see apply_kernel_anyk): two lanes, even / odd products, eight per
    // trip in the order 6,7 | 4,5 | 2,3 | 0,1, then pairs, lane 0 + lane 1
    double l0 = 0.0, l1 = 0.0;
    constexpr int full = (K / 8) * 8;
#pragma unroll
    for (int c = 0; c < full; c += 8) {
#pragma unroll
      for (int vv = 3; vv >= 0; --vv) {
        l0 = __dadd_rn(l0, __dmul_rn(w[c + 2 * vv], v[c + 2 * vv]));
        l1 = __dadd_rn(l1, __dmul_rn(w[c + 2 * vv + 1], v[c + 2 * vv + 1]));
      }
    }
#pragma unroll
    for (int c = full; c < K; c += 2) {
      l0 = __dadd_rn(l0, __dmul_rn(w[c], v[c]));
      l1 = __dadd_rn(l1, (c + 1 < K) ? __dmul_rn(w[c + 1 < K ? c + 1 : c], v[c + 1 < K ? c + 1 : c]) : 0.0);
    }
    return __dadd_rn(l0, l1);
  }
  double acc = __dmul_rn(w[0], v[0]);
#pragma unroll
  for (int kk = 1; kk < K; ++kk) acc = __dadd_rn(acc, __dmul_rn(w[kk], v[kk]));
  return acc;
}

constexpr unsigned long long kMarker = G2P_MASK_MARKER;

// MP: 0 no masks, 1 byte masks (PROPAGATE / FILL), 3 bit-packed masks (PROPAGATE_BITS), 4 MARKED source values
template <int MP>
__device__ __forceinline__ uint32_t src_masked(const ApplyArgs& a, int msg, int32_t i, double v) {
  if (MP == 1) return __ldg(a.src_mask + (int64_t)msg * a.src_stride + i) != 0;
  if (MP == 3) return (__ldg(a.src_mask + (int64_t)msg * a.sbits_stride + (i >> 3)) >> (i & 7)) & 1u;
  if (MP == 4) return (uint32_t)__double2hiint(v) == (uint32_t)(kMarker >> 32);  // upper word, as the windowed kernel
  return 0u;
}

template <int MP>
__device__ __forceinline__ void store_mask(const ApplyArgs& a, int msg, int64_t j, uint32_t mk) {
  if (!a.out_mask) return;
  if (MP == 1) a.out_mask[(int64_t)msg * a.out_stride + j] = (uint8_t)mk;
  if (MP >= 3 && mk)   // the launcher cleared the bit rows
    atomicOr(reinterpret_cast<uint32_t*>(a.out_mask) + (int64_t)msg * a.obits_stride + (j >> 5), 1u << (j & 31));
}

template <int K, int MP, int U>
__device__ __forceinline__ void apply_messages(const ApplyArgs& a, int msg, int64_t j0, bool has1,
                                               const int32_t (&i0)[K], const int32_t (&i1)[K],
                                               const double (&w0)[K], const double (&w1)[K]) {
  constexpr bool MASK = MP != 0;
  double v0[U][K], v1[U][K];
  uint32_t mk0[U], mk1[U];
  const uint32_t n = (uint32_t)a.n;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const double* __restrict__ s = a.src + (int64_t)(msg + u) * a.src_stride;
#pragma unroll
    for (int kk = 0; kk < K; ++kk) {
      v0[u][kk] = ((uint32_t)i0[kk] < n) ? __ldg(s + i0[kk]) : a.mv_out;
      v1[u][kk] = ((uint32_t)i1[kk] < n) ? __ldg(s + i1[kk]) : a.mv_out;
    }
    if (MASK) {
      uint32_t m0 = 0, m1 = 0;
#pragma unroll
      for (int kk = 0; kk < K; ++kk) {
        if ((uint32_t)i0[kk] < n) m0 |= src_masked<MP>(a, msg + u, i0[kk], v0[u][kk]);
        if ((uint32_t)i1[kk] < n) m1 |= src_masked<MP>(a, msg + u, i1[kk], v1[u][kk]);
      }
      mk0[u] = m0;
      mk1[u] = m1;
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    double r0 = weighted<K>(w0, v0[u], a.pairwise);
    double r1 = weighted<K>(w1, v1[u], a.pairwise);
    if (MASK) {  // masked outputs carry the target missing value (`ma.filled(result, mv_out)`)
      if (mk0[u]) r0 = a.mv_out;
      if (mk1[u]) r1 = a.mv_out;
    }
    double* o = a.out + (int64_t)(msg + u) * a.out_stride + j0;
    if (a.vec_store && has1) {
      __stcs(reinterpret_cast<double2*>(o), make_double2(r0, r1));
    } else {
      __stcs(o, r0);
      if (has1) __stcs(o + 1, r1);
    }
    if (MP >= 3) {
      // bit-packed out_mask: a warp owns 64 neighbouring targets starting at a multiple of 64; lanes 0, 4, .. 28
      // assemble one byte each from the two ballots (threads past the end of the table have left: they are not in
      // `act` and contribute zero bits)
      if (a.out_mask) {
        const unsigned act = __activemask();
        const uint32_t b0 = __ballot_sync(act, mk0[u] != 0), b1 = __ballot_sync(act, has1 && mk1[u] != 0);
        const int lane = threadIdx.x & 31;
        if ((lane & 3) == 0) {
          uint32_t n0 = (b0 >> lane) & 0xfu, n1 = (b1 >> lane) & 0xfu;
          n0 = (n0 | (n0 << 2)) & 0x33u;
          n0 = (n0 | (n0 << 1)) & 0x55u;
          n1 = (n1 | (n1 << 2)) & 0x33u;
          n1 = (n1 | (n1 << 1)) & 0x55u;
          a.out_mask[(int64_t)(msg + u) * a.obits_stride * 4 + (j0 >> 3)] = (uint8_t)(n0 | (n1 << 1));
        }
      }
    } else if (MASK) {
      store_mask<MP>(a, msg + u, j0, mk0[u]);
      if (has1) store_mask<MP>(a, msg + u, j0 + 1, mk1[u]);
    }
  }
}

template <int K, int MP>
__global__ void __launch_bounds__(kApplyThreads) apply_kernel(const ApplyArgs a) {
  for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int tt = (int)(tile % a.n_target_tiles);
    const int mc = (int)(tile / a.n_target_tiles);
    const int64_t j0 = ((int64_t)tt * kApplyThreads + threadIdx.x) * kTargetsPerThread;
    if (j0 >= a.m) continue;
    const bool has1 = (j0 + 1) < a.m;
    int32_t i0[K], i1[K];
    double w0[K], w1[K];
#pragma unroll
    for (int kk = 0; kk < K; ++kk) {
      const int64_t p = (int64_t)kk * a.m + j0;
      i0[kk] = __ldg(a.idx + p);
      i1[kk] = has1 ? __ldg(a.idx + p + 1) : (int32_t)a.n;
      if (K > 1) {
        w0[kk] = __ldg(a.coef + p);
        w1[kk] = has1 ? __ldg(a.coef + p + 1) : 0.0;
      } else {
        w0[kk] = 1.0;
        w1[kk] = 1.0;
      }
    }
    const int msg_begin = mc * a.msgs_per_chunk;
    const int msg_end = min(a.b, msg_begin + a.msgs_per_chunk);
    int msg = msg_begin;
    constexpr int MU = K > 4 ? 1 : kMsgUnroll;   // k = 11: the table of two targets already fills the registers
    for (; msg + MU <= msg_end; msg += MU)
      apply_messages<K, MP, MU>(a, msg, j0, has1, i0, i1, w0, w1);
    for (; msg < msg_end; ++msg) apply_messages<K, MP, 1>(a, msg, j0, has1, i0, i1, w0, w1);
  }
}

// generic k (2..16 other than 4): one target per thread, runtime loop; used for k=3/11 tables
template <int MP>
__global__ void __launch_bounds__(kApplyThreads) apply_kernel_anyk(const ApplyArgs a) {
  const int64_t total = a.m * (int64_t)a.b;
  const uint32_t n = (uint32_t)a.n;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int msg = (int)(t / a.m);
    const int64_t j = t - (int64_t)msg * a.m;
    const double* __restrict__ s = a.src + (int64_t)msg * a.src_stride;
    double acc = 0.0;
    uint32_t mk = 0;
    if (!a.pairwise) {
      for (int kk = 0; kk < a.k; ++kk) {
        const int32_t i = __ldg(a.idx + (int64_t)kk * a.m + j);
        const double w = __ldg(a.coef + (int64_t)kk * a.m + j);
        const bool in = (uint32_t)i < n;
        const double v = in ? __ldg(s + i) : a.mv_out;
        const double p = __dmul_rn(w, v);
        acc = (kk == 0) ? p : __dadd_rn(acc, p);
        if (MP != 0 && in) mk |= src_masked<MP>(a, msg, i, v);
      }
    } else {
      // G2P_SUM_PAIRWISE for any k: the order of numpy's buffered einsum (`sum_of_products_contig_two`, two SSE2 lanes
      // that take the even / odd products; eight products per trip in the order 6,7 | 4,5 | 2,3 | 0,1, then pairs, a
      // zero-filled last lane, lane 0 + lane 1) - what tables with float32 coefficients are summed with.  For k = 4
      // this is (p0 + p2) + (p1 + p3).
      double p[16];
      for (int kk = 0; kk < 16; ++kk) {
        p[kk] = 0.0;
        if (kk < a.k) {
          const int32_t i = __ldg(a.idx + (int64_t)kk * a.m + j);
          const double w = __ldg(a.coef + (int64_t)kk * a.m + j);
          const bool in = (uint32_t)i < n;
          const double v = in ? __ldg(s + i) : a.mv_out;
          p[kk] = __dmul_rn(w, v);
          if (MP != 0 && in) mk |= src_masked<MP>(a, msg, i, v);
        }
      }
      double l0 = 0.0, l1 = 0.0;
      int at = 0, left = a.k;
      for (; left >= 8; left -= 8, at += 8)
        for (int vv = 3; vv >= 0; --vv) {
          l0 = __dadd_rn(l0, p[at + 2 * vv]);
          l1 = __dadd_rn(l1, p[at + 2 * vv + 1]);
        }
      for (; left > 0; left -= 2, at += 2) {
        l0 = __dadd_rn(l0, p[at]);
        l1 = __dadd_rn(l1, left > 1 ? p[at + 1] : 0.0);
      }
      acc = __dadd_rn(l0, l1);
    }
    if (MP != 0 && mk) acc = a.mv_out;
    __stcs(a.out + (int64_t)msg * a.out_stride + j, acc);
    if (MP != 0) store_mask<MP>(a, msg, j, mk);
  }
}

// v = mask ? marker : v, in place (g2p_mark_masked)
template <bool BITS>
__global__ void mark_masked_kernel(double* __restrict__ src, const uint8_t* __restrict__ mask, int64_t n, int b,
                                   int64_t src_stride) {
  const int64_t total = n * (int64_t)b;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int msg = (int)(t / n);
    const int64_t i = t - (int64_t)msg * n;
    const uint32_t mk = BITS ? (mask[(int64_t)msg * (src_stride / 8) + (i >> 3)] >> (i & 7)) & 1u
                             : mask[(int64_t)msg * src_stride + i];
    if (mk) reinterpret_cast<unsigned long long*>(src)[(int64_t)msg * src_stride + i] = kMarker;
  }
}

template <typename Kern>
static int resident_ctas(Kern kern) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kApplyThreads, 0) != cudaSuccess || per_sm < 1)
    per_sm = 2;
  return per_sm * num_sms();
}

template <int K, int MP>
static int launch_apply(ApplyArgs a, cudaStream_t st) {
  static int grid_cap = 0;  // per (K, MP) instantiation
  if (grid_cap == 0) grid_cap = resident_ctas(apply_kernel<K, MP>);
  a.n_target_tiles = (int)ceil_div(a.m, (int64_t)kApplyThreads * kTargetsPerThread);
  // enough (target tile x message chunk) tiles for an even spread over the persistent grid,
  // but chunks as long as possible so the table is re-read (from L2) rarely
  int64_t want_chunks = ceil_div((int64_t)16 * grid_cap, a.n_target_tiles);
  int64_t max_chunks = ceil_div(a.b, kMsgUnroll * 2);
  int64_t n_chunks = want_chunks < 1 ? 1 : (want_chunks > max_chunks ? max_chunks : want_chunks);
  int per = (int)ceil_div(a.b, n_chunks);
  per = (int)(ceil_div(per, kMsgUnroll) * kMsgUnroll);
  a.msgs_per_chunk = per;
  n_chunks = ceil_div(a.b, per);
  a.n_tiles = (int64_t)a.n_target_tiles * n_chunks;
  int grid = (int)(a.n_tiles < grid_cap ? a.n_tiles : grid_cap);
  apply_kernel<K, MP><<<grid, kApplyThreads, 0, st>>>(a);
  G2P_LAUNCHED();
  return G2P_OK;
}

template <int MP>
static int launch_apply_anyk(ApplyArgs a, cudaStream_t st) {
  static int grid_cap = 0;
  if (grid_cap == 0) grid_cap = resident_ctas(apply_kernel_anyk<MP>);
  int64_t blocks = ceil_div(a.m * (int64_t)a.b, kApplyThreads);
  int grid = (int)(blocks < (int64_t)grid_cap * 4 ? blocks : (int64_t)grid_cap * 4);
  apply_kernel_anyk<MP><<<grid, kApplyThreads, 0, st>>>(a);
  G2P_LAUNCHED();
  return G2P_OK;
}

// ------------------------------------------------------------------ K6 pack / unpack
struct Rec16 {
  int64_t idx;
  double coef;
};

__global__ void pack_rec_kernel(const int32_t* __restrict__ idx, const double* __restrict__ coef, int64_t m, int k,
                                int coeff_f32, void* __restrict__ rec) {
  const int64_t total = m * k;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = r / k;
    const int kk = (int)(r - j * k);
    const int64_t i = idx[(int64_t)kk * m + j];
    const double w = coef[(int64_t)kk * m + j];
    if (!coeff_f32) {
      Rec16 o;
      o.idx = i;
      o.coef = w;
      reinterpret_cast<Rec16*>(rec)[r] = o;
    } else {  // packed 12-byte record {int64, float32}: three 4-byte words
      uint32_t* o = reinterpret_cast<uint32_t*>(rec) + 3 * r;
      o[0] = (uint32_t)(i & 0xffffffffLL);
      o[1] = (uint32_t)((uint64_t)i >> 32);
      o[2] = __float_as_uint((float)w);
    }
  }
}

__global__ void unpack_rec_kernel(const void* __restrict__ rec, int64_t m, int k, int coeff_f32,
                                  int32_t* __restrict__ idx, double* __restrict__ coef) {
  const int64_t total = m * k;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = r / k;
    const int kk = (int)(r - j * k);
    int64_t i;
    double w;
    if (!coeff_f32) {
      Rec16 in = reinterpret_cast<const Rec16*>(rec)[r];
      i = in.idx;
      w = in.coef;
    } else {
      const uint32_t* in = reinterpret_cast<const uint32_t*>(rec) + 3 * r;
      i = (int64_t)(((uint64_t)in[1] << 32) | in[0]);
      w = (double)__uint_as_float(in[2]);
    }
    // indexes are < 2^31 for every GRIB grid; anything else is mapped to -1 (= "no value")
    idx[(int64_t)kk * m + j] = (i >= 0 && i <= 0x7fffffffLL) ? (int32_t)i : -1;
    coef[(int64_t)kk * m + j] = w;
  }
}

}  // namespace g2p

using namespace g2p;

extern "C" {

int g2p_apply(const int32_t* idx, const double* coef, int64_t m, int k, const double* src, const uint8_t* src_mask,
              int64_t n, int b, int64_t src_stride, int64_t out_stride, double mv_out, int mask_policy, double* out,
              uint8_t* out_mask, void* stream) {
  if (!idx || !src || !out) return set_error(G2P_ERR_INVALID, "g2p_apply: null idx/src/out");
  if (k < 1 || k > 16) return set_error(G2P_ERR_INVALID, "g2p_apply: k=%d out of range 1..16", k);
  if (k > 1 && !coef) return set_error(G2P_ERR_INVALID, "g2p_apply: coef is required for k > 1");
  if (m < 0 || n < 0 || n > 0x7fffffffLL || b < 0) return set_error(G2P_ERR_INVALID, "g2p_apply: bad sizes");
  if (src_stride < n || out_stride < m) return set_error(G2P_ERR_INVALID, "g2p_apply: strides shorter than rows");
  const int pairwise = (mask_policy & G2P_SUM_PAIRWISE) != 0;
  mask_policy &= ~G2P_SUM_PAIRWISE;
  const bool mask = (mask_policy == G2P_MASK_PROPAGATE || mask_policy == G2P_MASK_FILL);   // byte masks
  const bool bits = mask_policy == G2P_MASK_PROPAGATE_BITS, marked = mask_policy == G2P_MASK_MARKED;
  if (!mask && !bits && !marked && mask_policy != G2P_MASK_AS_REFERENCE)
    return set_error(G2P_ERR_INVALID, "g2p_apply: unknown mask_policy %d", mask_policy);
  if ((mask || bits) && !src_mask) return set_error(G2P_ERR_INVALID, "g2p_apply: PROPAGATE / FILL need src_mask");
  if ((mask_policy == G2P_MASK_PROPAGATE || bits) && !out_mask)
    return set_error(G2P_ERR_INVALID, "g2p_apply: PROPAGATE needs out_mask");
  if (bits && (src_stride % 8) != 0) return set_error(G2P_ERR_INVALID, "g2p_apply: bit-packed masks need src_stride %% 8 == 0");
  if ((bits || marked) && out_mask && (reinterpret_cast<uintptr_t>(out_mask) & 3))
    return set_error(G2P_ERR_INVALID, "g2p_apply: a bit-packed out_mask must be 4-byte aligned");
  if (mask_policy == G2P_MASK_FILL) out_mask = nullptr;
  if (m == 0 || b == 0) return G2P_OK;
  ApplyArgs a{};
  a.idx = idx;
  a.coef = coef;
  a.src = src;
  a.src_mask = src_mask;
  a.out = out;
  a.out_mask = out_mask;
  a.m = m;
  a.n = n;
  a.src_stride = src_stride;
  a.out_stride = out_stride;
  a.sbits_stride = src_stride / 8;
  a.obits_stride = G2P_BITS_ROW_BYTES(out_stride) / 4;
  a.b = b;
  a.k = k;
  a.pairwise = pairwise;
  a.mv_out = mv_out;
  a.vec_store = ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && (out_stride % 2 == 0);
  cudaStream_t st = as_stream(stream);
  if ((bits || marked) && out_mask) {
    // k = 1 / 4: whole mask bytes come from warp ballots, only the padding bytes of the rows are cleared here;
    // other k: only masked outputs touch the rows (atomicOr) and everything else is this memset
    const size_t row_bytes = (size_t)a.obits_stride * 4, used = (size_t)((m + 7) / 8);
    if (k != 1 && k != 4 && k != 11) G2P_CUDA(cudaMemsetAsync(out_mask, 0, (size_t)b * row_bytes, st));
    else if (row_bytes > used) G2P_CUDA(cudaMemset2DAsync(out_mask + used, row_bytes, 0, row_bytes - used, (size_t)b, st));
  }
  const int mp = bits ? 3 : marked ? 4 : mask ? 1 : 0;
  if (k == 1)
    return mp == 3 ? launch_apply<1, 3>(a, st) : mp == 4 ? launch_apply<1, 4>(a, st) : mp == 1 ? launch_apply<1, 1>(a, st) : launch_apply<1, 0>(a, st);
  if (k == 4)
    return mp == 3 ? launch_apply<4, 3>(a, st) : mp == 4 ? launch_apply<4, 4>(a, st) : mp == 1 ? launch_apply<4, 1>(a, st) : launch_apply<4, 0>(a, st);
  // k = 11 (mode adw / cdd tables): the batched kernel keeps the 11 indexes and weights of its targets in registers and
  // reads them once per chunk of messages (the per-message kernel below re-reads 132 B of table per output value)
  static const bool anyk_only = getenv("G2P_APPLY_ANYK") != nullptr;
  if (k == 11 && !anyk_only)
    return mp == 3 ? launch_apply<11, 3>(a, st) : mp == 4 ? launch_apply<11, 4>(a, st) : mp == 1 ? launch_apply<11, 1>(a, st) : launch_apply<11, 0>(a, st);
  return mp == 3 ? launch_apply_anyk<3>(a, st) : mp == 4 ? launch_apply_anyk<4>(a, st) : mp == 1 ? launch_apply_anyk<1>(a, st) : launch_apply_anyk<0>(a, st);
}

int g2p_mark_masked(double* src, const uint8_t* src_mask, int mask_bits, int64_t n, int b, int64_t src_stride, void* stream) {
  if (!src || !src_mask) return set_error(G2P_ERR_INVALID, "g2p_mark_masked: null src / src_mask");
  if (n < 0 || b < 0 || src_stride < n) return set_error(G2P_ERR_INVALID, "g2p_mark_masked: bad sizes");
  if (mask_bits && (src_stride % 8) != 0) return set_error(G2P_ERR_INVALID, "g2p_mark_masked: bit-packed masks need src_stride %% 8 == 0");
  if (n == 0 || b == 0) return G2P_OK;
  const int64_t blocks = ceil_div(n * (int64_t)b, 256);
  const int grid = (int)(blocks < (int64_t)num_sms() * 16 ? blocks : (int64_t)num_sms() * 16);
  if (mask_bits) mark_masked_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(src, src_mask, n, b, src_stride);
  else mark_masked_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(src, src_mask, n, b, src_stride);
  G2P_LAUNCHED();
  return G2P_OK;
}

int g2p_table_pack_rec(const int32_t* idx, const double* coef, int64_t m, int k, int coeff_dtype, void* rec_out,
                       void* stream) {
  if (!idx || !coef || !rec_out) return set_error(G2P_ERR_INVALID, "g2p_table_pack_rec: null pointer");
  if (k < 1 || k > 16 || m < 0) return set_error(G2P_ERR_INVALID, "g2p_table_pack_rec: bad m/k");
  if (m == 0) return G2P_OK;
  int64_t blocks = ceil_div(m * k, 256);
  int grid = (int)(blocks < (int64_t)num_sms() * 16 ? blocks : (int64_t)num_sms() * 16);
  pack_rec_kernel<<<grid, 256, 0, as_stream(stream)>>>(idx, coef, m, k, coeff_dtype == G2P_F32, rec_out);
  G2P_LAUNCHED();
  return G2P_OK;
}

int g2p_table_unpack_rec(const void* rec, int64_t m, int k, int coeff_dtype, int32_t* idx_out, double* coef_out,
                         void* stream) {
  if (!rec || !idx_out || !coef_out) return set_error(G2P_ERR_INVALID, "g2p_table_unpack_rec: null pointer");
  if (k < 1 || k > 16 || m < 0) return set_error(G2P_ERR_INVALID, "g2p_table_unpack_rec: bad m/k");
  if (m == 0) return G2P_OK;
  int64_t blocks = ceil_div(m * k, 256);
  int grid = (int)(blocks < (int64_t)num_sms() * 16 ? blocks : (int64_t)num_sms() * 16);
  unpack_rec_kernel<<<grid, 256, 0, as_stream(stream)>>>(rec, m, k, coeff_dtype == G2P_F32, idx_out, coef_out);
  G2P_LAUNCHED();
  return G2P_OK;
}

}  // extern "C"
