// K8a: the source-grid manipulation stages that precede the intertable apply, fused into one
// gather-and-compute kernel: unit conversion -> step aggregation -> cut-off of negatives.
//
// Replaces, for all output messages of a run in one launch,
//   Converter.convert           where(x != mv, f(x), mv)              (manipulation/conversion.py:21,33-42)
//   Aggregator._accumulation    ((0.0 + V[t]) - V[t-D]) * unit / D    (manipulation/aggregator.py:72-153)
//   Aggregator._average         (0.0 + V[i1] + V[i2] + ...) / count   (:155-236), one add per hourly iterator;
//                               with `@halfweights` (0.0 + V[i1]*dt/2 + V[i2]*dt + ... + V[ik]*dt/2) / sum(dt)
//   Aggregator._instantaneous   0.0 + V[t]                            (:244-279)
//   Converter.cut_off_negative  where(v < 0, 0, v)                    (conversion.py:44-52)
// in the reference's evaluation order (numexpr evaluates left to right in float64; every operation below
// is a separate IEEE multiply / add / divide, no FMA contraction).  The controller's stage order is
// convert -> aggregate -> cut-off, all on the source grid, before interpolation (controller.py:106-132).
//
// With a `touched` list (the sorted source indexes an intertable references) the kernel writes COMPACT
// fields: out[o][c] = manip(V[..][touched[c]]), so only the 4 % of an O1280 field that the EFAS table
// reads is ever converted, aggregated and handed to the apply kernel (whose table is remapped to the
// compact numbering).  Without it the kernel is the plain elementwise stage on the whole grid.
#include <cmath>
#include <cstdlib>
#include <vector>

#include "g2p_common.cuh"

namespace g2p {

constexpr int kManipMaxIn = G2P_MANIP_MAX_IN;

struct ManipItem {       // device copy of g2p_manip_item
  int kind, n_in, mask_all, edge_mask;
  int in[kManipMaxIn];
  double unit_time, divisor, w_edge;
  double rcp;            // RN(1 / divisor), computed on the host
  int fast_div, pad;     // divisor is a normal number of moderate size: div_by may use rcp
};

struct ManipArgs {
  const double* src;
  const uint8_t* src_mask;
  const int32_t* touched;
  double* out;
  uint8_t* out_mask;
  int64_t n, nc, src_stride, out_stride;
  int b_in, n_out;
  int op1, op2;
  double c1, c2, mv_grib;
  int cut_off;
};

__device__ __forceinline__ double conv_op(int op, double c, double x) {
  switch (op) {
    case G2P_OP_MUL: return __dmul_rn(x, c);
    case G2P_OP_ADD: return __dadd_rn(x, c);
    case G2P_OP_DIV: return __ddiv_rn(x, c);
    default: return x;
  }
}

// x / d for a divisor that is the same for a whole message (the aggregation step, the count of an average), bit for
// bit the IEEE quotient.  __ddiv_rn is a long subroutine on this path - and a zero numerator (dry cells of an accumulated
// precipitation field: most of them) sends it through its slow branch.  Here: q = fma(r, y, q0) with y = RN(1/d),
// q0 = RN(x*y), r = x - q0*d (exact in an FMA) is the correctly rounded quotient in all but rare cases (Markstein), and
// instead of relying on that the candidate is CHECKED: its exact residual r1 = x - q*d must be strictly smaller than half
// an ulp of q times |d| (then q is the unique nearest double; powers of two, whose lower neighbour is closer, and
// results outside a safe exponent range are left to __ddiv_rn, like NaN and infinity, which fail the comparison).
__device__ __noinline__ double div_slow(double x, double d) { return __ddiv_rn(x, d); }   // rare: kept out of line

__device__ __forceinline__ double div_by(double x, double d, double y, int fast) {
  if (!fast) return div_slow(x, d);
  const double q0 = __dmul_rn(x, y);
  if (x == 0.0) return q0;                       // +-0 / d: the sign is the product's
  const double r0 = __fma_rn(-q0, d, x);
  const double q = __fma_rn(r0, y, q0);
  const double r1 = __fma_rn(-q, d, x);
  const int hi = __double2hiint(q);
  const int ex = (hi >> 20) & 0x7ff;
  const bool pow2 = ((hi & 0xfffff) | __double2loint(q)) == 0;
  const double half_ulp = __hiloint2double((ex - 53) << 20, 0);      // 2^(e - 53) for ex in the range tested below
  if (ex >= 200 && ex <= 1800 && !pow2 && fabs(r1) < __dmul_rn(fabs(d), half_ulp)) return q;
  return div_slow(x, d);
}

// where(x != mv, f(x), mv)  (conversion.py:21).  CONV: the conversion known at compile time - 0 identity, 1 `x*c`,
// 2 `x+c` (the one-step functions of parameters.json: `x*1000`, `x-273.15`, ...) - or 3: read the ops from the arguments
template <int CONV>
__device__ __forceinline__ double convert(const ManipArgs& a, double x) {
  if (CONV == 0) return x;
  if (CONV == 1) return x == a.mv_grib ? a.mv_grib : __dmul_rn(x, a.c1);
  if (CONV == 2) return x == a.mv_grib ? a.mv_grib : __dadd_rn(x, a.c1);
  if (a.op1 == G2P_OP_NONE) return x;  // 'x=x' is the identity: no guard either (conversion.py:34-35)
  if (x == a.mv_grib) return a.mv_grib;
  return conv_op(a.op2, a.c2, conv_op(a.op1, a.c1, x));
}


// value and mask of one output message at source point i.  KIND: the item kind when the whole launch has one kind
// (compile-time: the other branches disappear), -1 = read it from the item
template <int KIND, int CONV>
__device__ __forceinline__ double manip_point(const ManipArgs& a, const ManipItem& it, int64_t i, uint8_t& mk) {
  double res;
  const int kind = KIND >= 0 ? KIND : it.kind;
  if (kind == G2P_AGG_ACCUM) {
    // v_iter = 0.0 + V[t]; ((v_iter - V[t-D]) * unit_time) / D   (aggregator.py:142-147)
    const double cur = __dadd_rn(0.0, convert<CONV>(a, __ldg(a.src + (int64_t)it.in[0] * a.src_stride + i)));
    const double prev = it.in[1] >= 0 ? convert<CONV>(a, __ldg(a.src + (int64_t)it.in[1] * a.src_stride + i)) : 0.0;
    res = div_by(__dmul_rn(__dsub_rn(cur, prev), it.unit_time), it.divisor, it.rcp, it.fast_div);
  } else if (kind == G2P_AGG_AVERAGE) {
    // temp_sum = 0.0; temp_sum += V[next available step >= hour] per hourly iterator; / count (:209-230)
    double s = 0.0;
    for (int q = 0; q < it.n_in; ++q) s = __dadd_rn(s, convert<CONV>(a, __ldg(a.src + (int64_t)it.in[q] * a.src_stride + i)));
    res = div_by(s, it.divisor, it.rcp, it.fast_div);
  } else if (kind == G2P_AGG_WAVERAGE) {
    // `@halfweights`: v_ma = V[h] * dt (dt/2 for the first and last hour of the window); temp_sum += v_ma;
    // / sum of the weights (:186-207, :230).  V[h] * dt on a MaskedArray leaves the masked elements unmultiplied.
    double s = 0.0;
    for (int q = 0; q < it.n_in; ++q) {
      const int64_t at = (int64_t)it.in[q] * a.src_stride + i;
      const double x = convert<CONV>(a, __ldg(a.src + at));
      const bool edge = (q == 0 && (it.edge_mask & 1)) || (q == it.n_in - 1 && (it.edge_mask & 2));
      const bool masked = a.src_mask && __ldg(a.src_mask + at) != 0;
      s = __dadd_rn(s, masked ? x : __dmul_rn(x, edge ? it.w_edge : it.unit_time));
    }
    res = div_by(s, it.divisor, it.rcp, it.fast_div);
  } else if (kind == G2P_AGG_COPY) {
    res = convert<CONV>(a, a.src[(int64_t)it.in[0] * a.src_stride + i]);
  } else {
    // res = zeros; res += V[t]   (:264-276); in[0] < 0: step 0 absent -> zeros
    res = it.in[0] >= 0 ? __dadd_rn(0.0, convert<CONV>(a, __ldg(a.src + (int64_t)it.in[0] * a.src_stride + i))) : 0.0;
  }
  if (a.cut_off && res < 0.0) res = 0.0;  // where(v < 0, 0, v): NaN stays NaN
  mk = 0;
  if (a.src_mask) {
    if (it.mask_all < 0) {
      // what _average really returns: `numeric.get_masks(v_ord.values())` (aggregator.py:232) hands get_masks ONE
      // argument, the dict view, which is no MaskedArray - the mask is None and the averages come out unmasked
    } else if (it.mask_all == 1) {  // the OR of ALL inputs of the run (what the comment at aggregator.py:231 intends)
      for (int q = 0; q < a.b_in; ++q) mk |= __ldg(a.src_mask + (int64_t)q * a.src_stride + i);
    } else if (kind == G2P_AGG_ACCUM) {
      // the current step is copied into a fresh ndarray (`v_iter_ma += V[t]`, :142-143) and loses its
      // mask there: only the mask of V[t-D] survives (:148)
      if (it.in[1] >= 0) mk = __ldg(a.src_mask + (int64_t)it.in[1] * a.src_stride + i);
    } else if (kind == G2P_AGG_AVERAGE || kind == G2P_AGG_WAVERAGE) {
      for (int q = 0; q < it.n_in; ++q) mk |= __ldg(a.src_mask + (int64_t)it.in[q] * a.src_stride + i);
    } else if (kind == G2P_AGG_COPY) {
      mk = a.src_mask[(int64_t)it.in[0] * a.src_stride + i];
    }                          // PICK: `res_inst += V[t]` on a plain ndarray drops the mask (:264-276)
  }
  return res;
}

// PT points per thread: four independent gather chains in flight for the aggregations from device memory (+7 %); one
// point per thread for the plain compaction, whose reads go over PCIe when the messages lie in pinned host memory and
// want as many requests in flight as possible (4 per thread: -6 % end to end)
template <int kManipPerThread, int KIND, int CONV>
__device__ __forceinline__ void manip_chunk(const ManipArgs& a, const ManipItem& it, int o, int chunk) {
  // a warp owns 32 * PT consecutive points; lane l takes points l, l + 32, ...: every load and store instruction of the
  // warp covers 32 consecutive values (the touched source indexes of neighbouring points are consecutive too)
  const int lane = threadIdx.x & 31;
  const int64_t c0 = ((int64_t)chunk * blockDim.x + (threadIdx.x - lane)) * kManipPerThread + lane;
  if (c0 - lane >= a.nc) return;
  int64_t idx[kManipPerThread];
  double res[kManipPerThread];
  uint8_t mk[kManipPerThread];
#pragma unroll
  for (int u = 0; u < kManipPerThread; ++u) {
    const int64_t c = c0 + 32 * u < a.nc ? c0 + 32 * u : a.nc - 1;
    idx[u] = a.touched ? (int64_t)__ldg(a.touched + c) : c;
  }
#pragma unroll
  for (int u = 0; u < kManipPerThread; ++u) res[u] = manip_point<KIND, CONV>(a, it, idx[u], mk[u]);
#pragma unroll
  for (int u = 0; u < kManipPerThread; ++u) {
    const int64_t c = c0 + 32 * u;
    if (c >= a.nc) break;
    // no out_mask: a masked output carries the marker NaN instead (the MARKED input form of the apply kernels)
    if (a.src_mask && !a.out_mask && mk[u]) res[u] = __longlong_as_double((long long)G2P_MASK_MARKER);
    a.out[(int64_t)o * a.out_stride + c] = res[u];
    if (a.out_mask) a.out_mask[(int64_t)o * a.out_stride + c] = mk[u] ? 1 : 0;
  }
}

constexpr int kManipInline = 128;  // items passed by value in the kernel parameters (no upload, no sync; 21 KB of the 32 KB
                                   // parameter space): the 90 outputs of the eue_r24-style pipeline are one launch
struct ManipInline {
  ManipItem items[kManipInline];
};

// grid.y = output message of the launch; grid.x strides over the chunks of 256 * PT points (one CTA per chunk for the
// aggregations; the compaction over PCIe caps grid.x and loops)
template <int PT, int KIND, int CONV>
__global__ void __launch_bounds__(256) manip_kernel(const ManipArgs a, const __grid_constant__ ManipInline in, int o_base,
                                                    int n_chunks) {
  const int o = blockIdx.y;
  for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x)
    manip_chunk<PT, KIND, CONV>(a, in.items[o], o_base + o, chunk);
}

// V[a] + (V[b] - V[a]) * frac : linear-in-time synthesis of a missing step (aggregator.py:105,128)
__global__ void time_interp_kernel(const double* __restrict__ va, const double* __restrict__ vb, double frac, int64_t n,
                                   double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __dadd_rn(va[i], __dmul_rn(__dsub_rn(vb[i], va[i]), frac));
}

// per-target terms of Corrector.correct (correction.py:46,74-81): dem_term = dem * coeff, valid = both present
__global__ void correction_prepare_kernel(const void* __restrict__ dem, int dem_f32, double dem_mv,
                                          const double* __restrict__ gem, double gem_mv, double coeff, int64_t m,
                                          double* __restrict__ dem_term, uint8_t* __restrict__ valid) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (int64_t)gridDim.x * blockDim.x) {
    // a float32 dem map meets the float64 constant in numexpr: promoted to double before the product
    const double d = dem_f32 ? (double)reinterpret_cast<const float*>(dem)[j] : reinterpret_cast<const double*>(dem)[j];
    dem_term[j] = __dmul_rn(d, coeff);
    valid[j] = (d != dem_mv && gem[j] != gem_mv) ? 1 : 0;
  }
}

// standalone Corrector.correct on [b][m] outputs (used after the generic apply kernel); masked outputs stay as they are
__global__ void correct_kernel(double* __restrict__ v, const uint8_t* __restrict__ mask, int64_t m, int b, int64_t stride,
                               const double* __restrict__ gem, const double* __restrict__ dem_term,
                               const uint8_t* __restrict__ valid, double mv, int mode) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (int64_t)gridDim.x * blockDim.x) {
    const double g = gem[j], dt = dem_term[j];
    const bool ok = valid[j] != 0;
    for (int q = 0; q < b; ++q) {
      const int64_t at = (int64_t)q * stride + j;
      if (mask && mask[at]) continue;
      const double p = v[at];
      v[at] = !(ok && p != mv) ? mv : mode == 0 ? __dsub_rn(__dadd_rn(p, g), dt) : __dmul_rn(__ddiv_rn(p, g), dt);
    }
  }
}

// writers' post-processing in place (see g2p_writer_post in g2p.h)
__global__ void writer_post_kernel(double* __restrict__ v, const uint8_t* __restrict__ mask, int64_t m, int b, int64_t stride,
                                   const uint8_t* __restrict__ clone, double mv_match, int use_match, double mv_write) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (int64_t)gridDim.x * blockDim.x) {
    const bool out_of_clone = clone && clone[j] != 0;
    for (int q = 0; q < b; ++q) {
      const int64_t at = (int64_t)q * stride + j;
      const double p = v[at];
      if (out_of_clone || (mask && mask[at]) || p != p || (use_match && p == mv_match)) v[at] = mv_write;
    }
  }
}

}  // namespace g2p

using namespace g2p;

extern "C" {

int g2p_manip(const g2p_manip_desc* desc, const double* src, const uint8_t* src_mask, int64_t n, int b_in,
              int64_t src_stride, const int32_t* touched, int64_t nc, int64_t out_stride, double* out, uint8_t* out_mask,
              void* stream) {
  if (!desc || !src || !out) return set_error(G2P_ERR_INVALID, "g2p_manip: null desc/src/out");
  if (desc->n_out < 0 || (desc->n_out > 0 && !desc->items)) return set_error(G2P_ERR_INVALID, "g2p_manip: bad items");
  if (n < 0 || b_in < 1 || src_stride < n) return set_error(G2P_ERR_INVALID, "g2p_manip: bad sizes");
  if (!touched) nc = n;
  if (nc < 0 || out_stride < nc) return set_error(G2P_ERR_INVALID, "g2p_manip: out_stride shorter than the output rows");
  for (int op : {desc->conv_op1, desc->conv_op2})
    if (op < G2P_OP_NONE || op > G2P_OP_DIV) return set_error(G2P_ERR_INVALID, "g2p_manip: unknown conversion op %d", op);
  if (desc->n_out == 0 || nc == 0) return G2P_OK;
  if (desc->n_out > 65535) return set_error(G2P_ERR_UNSUPPORTED, "g2p_manip: more than 65535 output messages per call");
  std::vector<ManipItem> items(desc->n_out);
  for (int o = 0; o < desc->n_out; ++o) {
    const g2p_manip_item& s = desc->items[o];
    if (s.kind < G2P_AGG_PICK || s.kind > G2P_AGG_WAVERAGE) return set_error(G2P_ERR_INVALID, "g2p_manip: item %d: unknown kind %d", o, s.kind);
    const int need = s.kind == G2P_AGG_ACCUM ? 2 : ((s.kind == G2P_AGG_PICK || s.kind == G2P_AGG_COPY) ? 1 : s.n_in);
    if (s.n_in != need || need < (s.kind == G2P_AGG_WAVERAGE ? 0 : 1) || need > kManipMaxIn) return set_error(G2P_ERR_INVALID, "g2p_manip: item %d: n_in = %d", o, s.n_in);
    ManipItem& d = items[o];
    d.kind = s.kind;
    d.n_in = s.n_in;
    d.mask_all = s.mask_all;
    d.edge_mask = s.edge_mask;
    d.w_edge = s.w_edge;
    for (int q = 0; q < kManipMaxIn; ++q) {
      d.in[q] = q < s.n_in ? s.in[q] : -1;
      if (q < s.n_in && (s.in[q] >= b_in || (s.in[q] < 0 && !(s.kind == G2P_AGG_ACCUM && q == 1) && s.kind != G2P_AGG_PICK)))
        return set_error(G2P_ERR_INVALID, "g2p_manip: item %d references message %d of %d", o, s.in[q], b_in);
    }
    d.unit_time = s.unit_time;
    d.divisor = s.divisor;
    d.rcp = 1.0 / s.divisor;
    const double ad = fabs(s.divisor);
    d.fast_div = (std::isfinite(s.divisor) && ad >= 0x1p-100 && ad <= 0x1p100) ? 1 : 0;
  }
  cudaStream_t st = as_stream(stream);
  ManipArgs a{};
  a.src = src;
  a.src_mask = src_mask;
  a.touched = touched;
  a.out = out;
  a.out_mask = out_mask;
  a.n = n;
  a.nc = nc;
  a.src_stride = src_stride;
  a.out_stride = out_stride;
  a.b_in = b_in;
  a.n_out = desc->n_out;
  a.op1 = desc->conv_op1;
  a.op2 = desc->conv_op2;
  a.c1 = desc->conv_c1;
  a.c2 = desc->conv_c2;
  a.mv_grib = desc->mv_grib;
  a.cut_off = desc->cut_off_negative;
  bool all_copy = true, all_accum = true, all_pick = true;
  for (const ManipItem& it : items) {
    all_copy = all_copy && it.kind == G2P_AGG_COPY;
    all_accum = all_accum && it.kind == G2P_AGG_ACCUM;
    all_pick = all_pick && it.kind == G2P_AGG_PICK;
  }
  // points per thread: two independent gather chains for the aggregations (measured on the 90-output pipeline pass: 1:
  // 0.161 ms, 2: 0.140, 4: 0.179); one for the plain compaction, whose reads go over PCIe when the messages lie in pinned
  // host memory and want as many requests in flight as possible
  const int per_thread = all_copy ? 1 : 2;
  const int64_t n_chunks = ceil_div(nc, 256 * (int64_t)per_thread);
  for (int o0 = 0; o0 < desc->n_out; o0 += kManipInline) {
    const int cnt = std::min(kManipInline, desc->n_out - o0);
    ManipInline in{};
    for (int o = 0; o < cnt; ++o) in.items[o] = items[o0 + o];
    // the compaction over PCIe: a modest grid that strides through the points measured best (5 100 messages/s end to
    // end against 4 880 with one CTA per chunk); the aggregations: one CTA per chunk (a grid capped at 16 CTAs per SM
    // measured 0.265 against 0.232 ms for the 90-output pass)
    const int64_t cap = all_copy ? std::max<int64_t>(1, (int64_t)num_sms() * 16 / cnt) : 0x7fffffffLL;
    const dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(n_chunks, cap)), (unsigned)cnt);
    // kernel variant: item kind x conversion known at launch (the branches and the unused division code disappear)
    const int conv = a.op1 == G2P_OP_NONE ? 0 : (a.op2 == G2P_OP_NONE && a.op1 == G2P_OP_MUL) ? 1
                     : (a.op2 == G2P_OP_NONE && a.op1 == G2P_OP_ADD) ? 2 : 3;
#define G2P_MANIP_LAUNCH(PT, KIND)                                                                  \
  switch (conv) {                                                                                   \
    case 0: manip_kernel<PT, KIND, 0><<<grid, 256, 0, st>>>(a, in, o0, (int)n_chunks); break;       \
    case 1: manip_kernel<PT, KIND, 1><<<grid, 256, 0, st>>>(a, in, o0, (int)n_chunks); break;       \
    case 2: manip_kernel<PT, KIND, 2><<<grid, 256, 0, st>>>(a, in, o0, (int)n_chunks); break;       \
    default: manip_kernel<PT, KIND, 3><<<grid, 256, 0, st>>>(a, in, o0, (int)n_chunks); break;      \
  }
    if (all_copy) {
      G2P_MANIP_LAUNCH(1, G2P_AGG_COPY)
    } else if (all_accum) {
      G2P_MANIP_LAUNCH(2, G2P_AGG_ACCUM)
    } else if (all_pick) {
      G2P_MANIP_LAUNCH(2, G2P_AGG_PICK)
    } else {
      manip_kernel<2, -1, 3><<<grid, 256, 0, st>>>(a, in, o0, (int)n_chunks);
    }
#undef G2P_MANIP_LAUNCH
    G2P_LAUNCHED();
  }
  return G2P_OK;
}

int g2p_time_interp(const double* va, const double* vb, double frac, int64_t n, double* out, void* stream) {
  if (!va || !vb || !out || n < 0) return set_error(G2P_ERR_INVALID, "g2p_time_interp: bad arguments");
  if (n == 0) return G2P_OK;
  int64_t blocks = std::min<int64_t>(ceil_div(n, 256), (int64_t)num_sms() * 16);
  time_interp_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(va, vb, frac, n, out);
  G2P_LAUNCHED();
  return G2P_OK;
}

int g2p_correction_prepare(const void* dem, int dem_dtype, double dem_mv, const double* gem, double gem_mv, double coeff,
                           int64_t m, double* dem_term_out, uint8_t* valid_out, void* stream) {
  if (!dem || !gem || !dem_term_out || !valid_out || m < 0) return set_error(G2P_ERR_INVALID, "g2p_correction_prepare: bad arguments");
  if (dem_dtype != G2P_F64 && dem_dtype != G2P_F32) return set_error(G2P_ERR_INVALID, "g2p_correction_prepare: bad dtype");
  if (m == 0) return G2P_OK;
  int64_t blocks = std::min<int64_t>(ceil_div(m, 256), (int64_t)num_sms() * 16);
  correction_prepare_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(dem, dem_dtype == G2P_F32, dem_mv, gem, gem_mv,
                                                                            coeff, m, dem_term_out, valid_out);
  G2P_LAUNCHED();
  return G2P_OK;
}

int g2p_correct(double* values, const uint8_t* mask, int64_t m, int b, int64_t stride, const g2p_correction* corr,
                void* stream) {
  if (!values || !corr || !corr->gem || !corr->dem_term || !corr->valid || m < 0 || b < 0 || stride < m)
    return set_error(G2P_ERR_INVALID, "g2p_correct: bad arguments");
  if (corr->mode != 0 && corr->mode != 1) return set_error(G2P_ERR_INVALID, "g2p_correct: unknown correction mode %d", corr->mode);
  if (m == 0 || b == 0) return G2P_OK;
  int64_t blocks = std::min<int64_t>(ceil_div(m, 256), (int64_t)num_sms() * 16);
  correct_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(values, mask, m, b, stride, corr->gem, corr->dem_term,
                                                                 corr->valid, corr->mv, corr->mode);
  G2P_LAUNCHED();
  return G2P_OK;
}

int g2p_writer_post_apply(double* values, const uint8_t* mask, int64_t m, int b, int64_t stride, const g2p_writer_post* post,
                          void* stream) {
  if (!values || !post || m < 0 || b < 0 || stride < m) return set_error(G2P_ERR_INVALID, "g2p_writer_post_apply: bad arguments");
  if (post->out_f32) return set_error(G2P_ERR_UNSUPPORTED, "g2p_writer_post_apply: the in-place pass keeps float64 (out_f32 is an epilogue of g2p_apply_planned_post)");
  if (m == 0 || b == 0) return G2P_OK;
  int64_t blocks = std::min<int64_t>(ceil_div(m, 256), (int64_t)num_sms() * 16);
  writer_post_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(values, mask, m, b, stride, post->clone_mask,
                                                                     post->mv_match, post->use_match, post->mv_write);
  G2P_LAUNCHED();
  return G2P_OK;
}

}  // extern "C"
