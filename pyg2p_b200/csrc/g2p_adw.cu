// K5b  Angular-distance weighting (mode `adw`, Shepard 1968, k = 11) on sm_100a.
//
// Replaces the Shepard branch of ScipyInterpolation._build_weights_invdist
// (reference scipy_interpolation_lib.py:817-838, :953-1001, :1019-1021, :1036-1052) and its numba kernel
// adw_compute_weights_from_cutoff_distances (:322-381).  Input: the raw k-NN result of g2p_knn (k = 11 neighbour ids
// and distances per target); output: the normalised weights, in place of the distances.
//
// One thread owns one target: 11 distances, 11 neighbour coordinates (gathered once), the 11 x 10 angular terms and the
// normalisation all stay in registers.  Every operation is a separately rounded IEEE operation in the reference's
// order (numpy evaluates array expressions operator by operator; numba compiles without fast-math):
//   r'   = d(5) where d(4) > r_ref, r_ref where d(4) <= r_ref < d(10), else d(11)
//   s_j  = 1/d_j for d_j <= r'/3,  (27/(4 r')) * (d_j/r' - 1)^2 for r'/3 < d_j <= r',  else 0
//   t_i  = sum_{j != i} (1 - cos_ij) s_j / sum_{j != i} s_j   with cos_ij on the lat/lon plane, NaN -> 1
//          (both sums run sequentially over j: numpy reduces the Fortran-ordered result of `a[:, others]` that way)
//   w_i  = s_i^2 (1 + t_i),  weights = w / (((w0+w1)+(w2+w3)) + ((w4+w5)+(w6+w7)) + w8 + w9 + w10)   (numpy's pairwise sum)
//   d(1) <= 1e-10 -> [1, 0, ..., 0]
// float32 target maps (PCRaster REAL4): distances, r', s, t, w are float32 arrays in the reference; the arithmetic
// between the stores follows numba's / numpy's typing (float32 array with an int literal -> float64; array / array stays
// float32), restated per line below.  Bound: HBM (k*12 B in, k*8 B out per target, plus the neighbour coordinates from L2).
#include <float.h>

#include "g2p_common.cuh"

namespace g2p {

constexpr int kAdwK = 11;

// ---- reference radius: the exact sum of the 7th distances (double-double accumulation: the rounded result does not
// depend on the grid shape, the rank partition or the order)
struct DDsum {
  double hi, lo;
};
__device__ __forceinline__ void dd_acc(DDsum& a, double v) {
  const double s = __dadd_rn(a.hi, v);
  const double bb = __dsub_rn(s, a.hi);
  const double err = __dadd_rn(__dsub_rn(a.hi, __dsub_rn(s, bb)), __dsub_rn(v, bb));
  a.hi = s;
  a.lo = __dadd_rn(a.lo, err);
}
__device__ __forceinline__ void dd_merge(DDsum& a, double hi, double lo) {
  dd_acc(a, hi);
  a.lo = __dadd_rn(a.lo, lo);
}

__global__ void __launch_bounds__(256) column_sum_kernel(const double* __restrict__ col, int64_t m, int f32,
                                                         double* __restrict__ partial /* [grid][2] */) {
  DDsum acc{0.0, 0.0};
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < m; q += (int64_t)gridDim.x * blockDim.x) {
    double v = col[q];
    if (f32) v = (double)(float)v;
    dd_acc(acc, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double hi = __shfl_down_sync(0xffffffffu, acc.hi, o);
    const double lo = __shfl_down_sync(0xffffffffu, acc.lo, o);
    dd_merge(acc, hi, lo);
  }
  __shared__ double sh[8][2];
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    sh[warp][0] = acc.hi;
    sh[warp][1] = acc.lo;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    DDsum t{sh[0][0], sh[0][1]};
    for (int w = 1; w < 8; ++w) dd_merge(t, sh[w][0], sh[w][1]);
    partial[2 * blockIdx.x] = t.hi;
    partial[2 * blockIdx.x + 1] = t.lo;
  }
}

__device__ __forceinline__ double nan_to_num1(double v, double big) {
  if (v != v) return 1.0;                 // np.nan_to_num(..., nan=1): no other station in range (0/0)
  if (v > big) return big;                // +-inf -> the largest finite value of the array's dtype
  if (v < -big) return -big;
  return v;
}

template <bool F32>
__global__ void __launch_bounds__(128) adw_weights_kernel(int64_t m, double r_ref, const void* __restrict__ tlat,
                                                          const void* __restrict__ tlon, const double* __restrict__ slat,
                                                          const double* __restrict__ slon, int64_t n_src,
                                                          const int32_t* __restrict__ idx, double* __restrict__ coef) {
  constexpr int K = kAdwK;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < m; q += (int64_t)gridDim.x * blockDim.x) {
    double d[K], s[K], x[K], y[K], D[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      d[j] = coef[(int64_t)j * m + q];
      if (F32) d[j] = (double)(float)d[j];
    }
    // ---- cut-off radius r' (:343-355)
    double r = d[10];
    if (d[3] > r_ref)
      r = d[4];
    else if (d[9] > r_ref)
      r = F32 ? (double)(float)r_ref : r_ref;         // stored into the float32 array r
    // ---- s (:360-369)
    const double third = __ddiv_rn(r, 3.0);
#pragma unroll
    for (int j = 0; j < K; ++j) {
      double v = 0.0;
      if (d[j] <= third) {
        v = __ddiv_rn(1.0, d[j]);
      } else if (d[j] <= r) {
        // float32 maps: `distances / r_broadcasted` is float32 / float32, `- 1` and `** 2` are float64
        const double ratio = F32 ? (double)__fdiv_rn((float)d[j], (float)r) : __ddiv_rn(d[j], r);
        const double e = __dsub_rn(ratio, 1.0);
        v = __dmul_rn(__ddiv_rn(27.0, __dmul_rn(4.0, r)), __dmul_rn(e, e));
      }
      s[j] = F32 ? (double)(float)v : v;
    }
    // ---- neighbour offsets on the lat/lon plane (:955-958); x is the latitude difference, as in the reference
    const double la = F32 ? (double)((const float*)tlat)[q] : ((const double*)tlat)[q];
    const double lo = F32 ? (double)((const float*)tlon)[q] : ((const double*)tlon)[q];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      int64_t id = idx[(int64_t)j * m + q];
      if (id < 0 || id >= n_src) id = 0;               // cannot happen for a k-NN result with n_src >= 11 points
      x[j] = __dsub_rn(la, slat[id]);
      y[j] = __dsub_rn(lo, slon[id]);
      D[j] = __dsqrt_rn(__dadd_rn(__dmul_rn(x[j], x[j]), __dmul_rn(y[j], y[j])));
    }
    // ---- directional term, weights (:966-973, :988, :1001)
    double w[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      double num = 0.0, den = 0.0;
      float denf = 0.0f;
      bool first = true;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        if (j == i) continue;
        const double c = __ddiv_rn(__dadd_rn(__dmul_rn(x[i], x[j]), __dmul_rn(y[i], y[j])), __dmul_rn(D[i], D[j]));
        const double t = __dmul_rn(__dsub_rn(1.0, c), s[j]);
        if (first) {
          num = t;
          den = s[j];
          denf = (float)s[j];
          first = false;
        } else {
          num = __dadd_rn(num, t);
          if (F32)
            denf = __fadd_rn(denf, (float)s[j]);       // np.sum of the float32 array s[:, others]
          else
            den = __dadd_rn(den, s[j]);
        }
      }
      double wd = __ddiv_rn(num, F32 ? (double)denf : den);
      if (F32) {
        // stored into the float32 array weight_directional, then `w * (1 + wd)` in float32
        const float wdf = (float)nan_to_num1((double)(float)wd, (double)FLT_MAX);
        const float sf = (float)s[i];
        w[i] = (double)__fmul_rn(__fmul_rn(sf, sf), __fadd_rn(1.0f, wdf));
      } else {
        wd = nan_to_num1(wd, DBL_MAX);
        w[i] = __dmul_rn(__dmul_rn(s[i], s[i]), __dadd_rn(1.0, wd));
      }
    }
    // ---- normalise: numpy's pairwise sum of 11 contiguous values (:1019-1021)
    double out[K];
    if (F32) {
      float f[K];
#pragma unroll
      for (int j = 0; j < K; ++j) f[j] = (float)w[j];
      float sum = __fadd_rn(__fadd_rn(__fadd_rn(f[0], f[1]), __fadd_rn(f[2], f[3])),
                            __fadd_rn(__fadd_rn(f[4], f[5]), __fadd_rn(f[6], f[7])));
      sum = __fadd_rn(__fadd_rn(__fadd_rn(sum, f[8]), f[9]), f[10]);
#pragma unroll
      for (int j = 0; j < K; ++j) out[j] = (double)__fdiv_rn(f[j], sum);
    } else {
      double sum = __dadd_rn(__dadd_rn(__dadd_rn(w[0], w[1]), __dadd_rn(w[2], w[3])),
                             __dadd_rn(__dadd_rn(w[4], w[5]), __dadd_rn(w[6], w[7])));
      sum = __dadd_rn(__dadd_rn(__dadd_rn(sum, w[8]), w[9]), w[10]);
#pragma unroll
      for (int j = 0; j < K; ++j) out[j] = __ddiv_rn(w[j], sum);
    }
    // ---- a target on a station takes exactly that station (:1041-1050)
    const bool exact = F32 ? ((float)d[0] <= 1e-10f) : (d[0] <= 1e-10);
#pragma unroll
    for (int j = 0; j < K; ++j) coef[(int64_t)j * m + q] = exact ? (j == 0 ? 1.0 : 0.0) : out[j];
  }
}

}  // namespace g2p

using namespace g2p;

extern "C" {

int g2p_adw_distance_sum(const double* dist, int64_t m, int k, int column, int dtype, double* sum_out, void* stream) {
  if (!dist || !sum_out) return set_error(G2P_ERR_INVALID, "g2p_adw_distance_sum: null pointer");
  if (k < 1 || column < 0 || column >= k || m < 0) return set_error(G2P_ERR_INVALID, "g2p_adw_distance_sum: bad sizes");
  sum_out[0] = sum_out[1] = 0.0;
  if (m == 0) return G2P_OK;
  cudaStream_t st = as_stream(stream);
  const int grid = (int)std::min<int64_t>(ceil_div(m, 256), (int64_t)num_sms() * 8);
  double* partial = nullptr;
  G2P_CUDA(cudaMallocAsync(&partial, sizeof(double) * 2 * grid, st));
  column_sum_kernel<<<grid, 256, 0, st>>>(dist + (int64_t)column * m, m, dtype == G2P_F32 ? 1 : 0, partial);
  G2P_LAUNCHED();
  double* h = (double*)malloc(sizeof(double) * 2 * grid);
  if (!h) return set_error(G2P_ERR_NOMEM, "g2p_adw_distance_sum: host scratch");
  cudaError_t e = cudaMemcpyAsync(h, partial, sizeof(double) * 2 * grid, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFreeAsync(partial, st);
  if (e != cudaSuccess) {
    free(h);
    return cuda_fail(e, "g2p_adw_distance_sum", __FILE__, __LINE__);
  }
  // exact-ish host merge of the block partials (two-sum)
  double hi = 0.0, lo = 0.0;
  for (int b = 0; b < grid; ++b) {
    const double v = h[2 * b];
    const double s = hi + v;
    const double bb = s - hi;
    lo += (hi - (s - bb)) + (v - bb);
    lo += h[2 * b + 1];
    hi = s;
  }
  free(h);
  const double s = hi + lo;
  sum_out[0] = s;
  sum_out[1] = lo - (s - hi);
  return G2P_OK;
}

int g2p_weights_adw(int64_t m, int k, int dtype, double ref_radius, const void* tlat, const void* tlon, const double* slat,
                    const double* slon, int64_t n_src, const int32_t* idx, double* coef, void* stream) {
  if (!tlat || !tlon || !slat || !slon || !idx || !coef) return set_error(G2P_ERR_INVALID, "g2p_weights_adw: null pointer");
  if (k != kAdwK) return set_error(G2P_ERR_UNSUPPORTED, "g2p_weights_adw: the Shepard selection needs k = 11, got %d", k);
  if (dtype != G2P_F64 && dtype != G2P_F32) return set_error(G2P_ERR_INVALID, "g2p_weights_adw: bad dtype");
  if (m < 0 || n_src < 1) return set_error(G2P_ERR_INVALID, "g2p_weights_adw: bad sizes");
  if (m == 0) return G2P_OK;
  cudaStream_t st = as_stream(stream);
  const int grid = (int)std::min<int64_t>(ceil_div(m, 128), (int64_t)num_sms() * 32);
  if (dtype == G2P_F32)
    adw_weights_kernel<true><<<grid, 128, 0, st>>>(m, ref_radius, tlat, tlon, slat, slon, n_src, idx, coef);
  else
    adw_weights_kernel<false><<<grid, 128, 0, st>>>(m, ref_radius, tlat, tlon, slat, slon, n_src, idx, coef);
  G2P_LAUNCHED();
  return G2P_OK;
}

}  // extern "C"
