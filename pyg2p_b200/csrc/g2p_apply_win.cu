// K7w: windowed (shared-memory staged) batched intertable apply for sm_100a, plus the tile plan it needs.
//
// Same arithmetic and semantics as g2p_apply (g2p_apply.cu; reference
// src/pyg2p/main/interpolation/__init__.py:142-162), different data movement.  The generic kernel
// gathers straight from global memory: every warp-level gather touches 10-16 sectors, the L1
// wavefront queue - not HBM - is the limit (ncu: 15.7 sectors/request, 14 % of HBM peak).
// Here the targets are cut into 2-D tiles of 1024; the source points a tile references form a few
// contiguous index ranges (pieces of neighbouring latitude rings).  A plan, built once per table,
// lists those ranges per tile (16-element aligned) and rewrites the table's int32 indexes as
// uint16 byte offsets into the tile's window.  The apply kernel then, per (tile, message):
//   * all threads copy the ranges - values and mask bytes - into a shared-memory stage with 16-byte
//     cp.async (LDGSTS, L2-only), 2-4 stages in flight.  (A first version issued one cp.async.bulk
//     per range from one warp: the ranges are ~32 values long, 44 bulk copies per tile and message,
//     and UBLKCP takes its operands from uniform registers, so the issuing warp serialised them at
//     ~20 instructions each and every other warp waited at the barrier - see profiles/.)
//   * one message ahead of the gathers, a fold pass writes a NaN marker over every staged value
//     whose mask byte is set, so the inner loop has no mask gathers (a NaN result takes a rare path
//     that looks for the marker among its operands);
//   * all threads gather their k neighbours from shared memory - neighbouring lanes own neighbouring
//     targets, and the plan places the ranges (latitude rings) at offsets skewed by their measured
//     longitude alignment so that neighbouring rings fall into disjoint bank groups -, do the no-FMA
//     weighted sum (and optionally the Corrector formula) and write the outputs with streaming stores,
//     128 contiguous bytes per half-warp.
// Every source value is read from L2/HBM once per tile instead of once per referencing target, and
// index/weight registers are loaded once per (tile, message chunk).  Variants by template flag: mask
// policy, Corrector epilogue (CORR), source-grid stages fused into the fold pass (AGG).
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "g2p_common.cuh"

namespace g2p {

constexpr int kWinThreads = 256;
constexpr int kWinTPT = 4;                               // targets per thread: 2 adjacent pairs
constexpr int kTileTargets = kWinThreads * kWinTPT;      // 1024
constexpr int kMaxRanges = 96;                           // source index ranges per tile
constexpr int kMaxStages = 4;
constexpr int kBlockElems = 16;                          // range granularity: 16 values = 128 B, 16 mask bytes
constexpr int kPlanSpanBlocks = 1 << 17;                 // bitmap span per tile: 2 M source elements
constexpr int kMaxWindowElems = 8176;                    // 16 x 16-byte value chunks per thread and message; byte offsets fit uint16
constexpr uint16_t kLocalSentinel = 0xffff;

struct __align__(16) Range {
  int32_t src_start;   // first source element (multiple of 16)
  int32_t n_elem;      // multiple of 16
  int32_t smem_off;    // element offset of the values inside the stage window (skewed, even)
  int32_t mask_off;    // element offset of the mask bytes inside the stage's mask window (dense)
};

}  // namespace g2p

struct g2p_plan {
  int device;
  int64_t m, n, n_pad;
  int k;
  int64_t rows, cols;
  int tile_rows, tile_cols, tiles_r, tiles_c, n_tiles;
  int max_window;      // elements, over tiles
  int max_ranges;
  double mean_window;
  uint16_t* lidx;      // [k][m] byte offsets into the tile's value window (0xffff: no value)
  g2p::Range* ranges;  // [n_tiles][kMaxRanges]
  int32_t* tile_info;  // [n_tiles][4] = {n_ranges, value window (skewed), mask window (dense), 0}
  mutable void* item_scratch;       // device copy of the fused pipeline's item list (grown on demand)
  mutable size_t item_scratch_bytes;
  int max_mwindow;     // largest dense (mask) window
  int block_cols_log2; // half-warp block shape (TileGeom)
  int skew_mod, skew_step;
  double gather_wavefronts;  // mean shared-memory wavefronts per half-warp gather (1 = conflict-free)
};

namespace g2p {

// ------------------------------------------------------------------ tile geometry (shared by plan and apply)
struct TileGeom {
  int64_t rows, cols;
  int tile_rows, tile_cols, tiles_c, threads_per_row;
  int block_cols_log2;  // a half-warp owns a block of (16 >> block_cols_log2) rows x (1 << block_cols_log2) columns
};

// The threads of a tile form a grid of tile_rows x T (T = threads per row); thread (tr, x) owns the target
// columns x, x + T, x + 2T, x + 3T of tile row tr.  A half-warp (the unit a 64-bit shared-memory load is served
// in) owns a compact block of that grid - 1x16, 2x8 or 4x4 - so that its k-th neighbours are the same or
// adjacent source points of at most a few rings (broadcast / distinct banks in shared memory); the two
// half-warps of a warp sit side by side, so a warp stores 2 x 128 (2x8 blocks) contiguous bytes per target slot.
__device__ __forceinline__ void thread_targets(const TileGeom& g, int tile, int tid, int64_t& r, int64_t (&c)[kWinTPT]) {
  const int ty = tile / g.tiles_c, tx = tile - ty * g.tiles_c;
  const int hw = tid >> 4, l = tid & 15;
  const int bpr = g.threads_per_row >> g.block_cols_log2;  // blocks per row of blocks
  const int by = hw / bpr, bx = hw - by * bpr;
  const int tr = by * (16 >> g.block_cols_log2) + (l >> g.block_cols_log2);
  const int x = (bx << g.block_cols_log2) + (l & ((1 << g.block_cols_log2) - 1));
  r = (int64_t)ty * g.tile_rows + tr;
  const int64_t c0 = (int64_t)tx * g.tile_cols + x;
#pragma unroll
  for (int t = 0; t < kWinTPT; ++t) c[t] = c0 + (int64_t)t * g.threads_per_row;
}

// ------------------------------------------------------------------ plan kernel: one CTA per tile
struct PlanArgs {
  const int32_t* idx;
  int64_t m, n, n_pad;
  int k;
  TileGeom g;
  uint16_t* lidx;
  Range* ranges;
  int32_t* tile_info;
  int32_t* status;  // [0] = first failing reason (0 ok), [1] = max value window, [2] = max ranges, [3] = max mask window
  unsigned long long* total_window;
  unsigned long long* total_wavefronts;  // shared-memory wavefronts of all half-warp gathers of one message
  int swap01;       // 1: targets load neighbours 0 / 1 in either order, 2: and 2 / 3 (k = 4); 0: distance order
  int swap_sweeps;  // refinement passes of pick_pair_order
  int dry;          // only measure (no lidx / ranges written): used to choose the tile shape
  int skew_mod, skew_step;  // range i goes to bank group skew_step * (i % skew_mod)
};

// wavefronts of one 64-bit shared-memory load of a half-warp: distinct words on the busiest of the 16 8-byte banks
__device__ __forceinline__ int half_warp_wavefronts(unsigned word, int hw_base, int hl) {
  const unsigned half = 0xffffu << hw_base;
  const unsigned same = __match_any_sync(0xffffffffu, word) & half;
  const bool first = (__ffs(same) - 1) == hw_base + hl;           // no lower lane of the half-warp loads this word
  const unsigned firsts = __ballot_sync(0xffffffffu, first);
  const unsigned bank = __match_any_sync(0xffffffffu, word & 15u) & half;
  int cnt = __popc(bank & firsts);
  for (int o = 8; o > 0; o >>= 1) cnt = max(cnt, __shfl_xor_sync(0xffffffffu, cnt, o));
  return cnt;
}

// Every lane of a half-warp loads word A with one instruction and word B with another, and may exchange the two.
// Greedy over the 16 lanes (lane q takes the orientation that adds fewer wavefronts to the two loads given the
// choices of lanes 0..q-1), then `sweeps` passes that flip single lanes while the two loads get cheaper.  Returns
// whether this lane exchanges its pair; A and B come back in the chosen order.  Called by whole warps.
__device__ __forceinline__ bool pick_pair_order(unsigned& A, unsigned& B, bool can, int hw_base, int hl, int sweeps) {
  bool first_a = false, first_b = false, swapped = false;
  int cur_a = 0, cur_b = 0;
  auto hb = [&](bool pred) { return (__ballot_sync(0xffffffffu, pred) >> hw_base) & 0xffffu; };
  for (int q = 0; q < 16; ++q) {
    const unsigned x = __shfl_sync(0xffffffffu, A, hw_base + q), y = __shfl_sync(0xffffffffu, B, hw_base + q);
    const int cq = __shfl_sync(0xffffffffu, (int)can, hw_base + q);
    const bool dec = hl < q;   // this lane has chosen
    // distinct words already on the candidate's bank (+1 for the candidate unless some lane loads the same word)
    const int n_ax = __popc(hb(dec && first_a && ((A ^ x) & 15u) == 0)), n_ay = __popc(hb(dec && first_a && ((A ^ y) & 15u) == 0));
    const int n_bx = __popc(hb(dec && first_b && ((B ^ x) & 15u) == 0)), n_by = __popc(hb(dec && first_b && ((B ^ y) & 15u) == 0));
    const bool p_ax = hb(dec && A == x) != 0, p_ay = hb(dec && A == y) != 0;
    const bool p_bx = hb(dec && B == x) != 0, p_by = hb(dec && B == y) != 0;
    const int c_ax = p_ax ? 0 : n_ax + 1, c_ay = p_ay ? 0 : n_ay + 1, c_bx = p_bx ? 0 : n_bx + 1, c_by = p_by ? 0 : n_by + 1;
    const int keep = max(cur_a, c_ax) + max(cur_b, c_by), swap = max(cur_a, c_ay) + max(cur_b, c_bx);
    const bool sw = cq && swap < keep;   // the same in all lanes of the half-warp
    if (hl == q) {
      if (sw) {
        const unsigned tmp = A;
        A = B;
        B = tmp;
        swapped = true;
      }
      first_a = !(sw ? p_ay : p_ax);
      first_b = !(sw ? p_bx : p_by);
    }
    cur_a = max(cur_a, sw ? c_ay : c_ax);
    cur_b = max(cur_b, sw ? c_bx : c_by);
  }
  if (sweeps > 0) {
    int best = half_warp_wavefronts(A, hw_base, hl) + half_warp_wavefronts(B, hw_base, hl);
    for (int sweep = 0; sweep < sweeps; ++sweep) {
      bool improved = false;   // the same in all lanes of the half-warp
      for (int q = 0; q < 16; ++q) {
        const bool me = hl == q;
        const int cq = __shfl_sync(0xffffffffu, (int)can, hw_base + q);
        const unsigned a2 = me ? B : A, b2 = me ? A : B;
        const int c = half_warp_wavefronts(a2, hw_base, hl) + half_warp_wavefronts(b2, hw_base, hl);
        if (cq && c < best) {
          best = c;
          improved = true;
          if (me) swapped = !swapped;
          A = a2;
          B = b2;
        }
      }
      // the two half-warps of a warp run the collectives together: no early exit unless both are done
      if (!__any_sync(0xffffffffu, improved)) break;
    }
  }
  return swapped;
}

__global__ void __launch_bounds__(kWinThreads) plan_kernel(const PlanArgs a) {
  extern __shared__ uint32_t s_bits[];  // kPlanSpanBlocks / 32 words
  __shared__ int s_min, s_max, s_nr, s_fail, s_mwin;
  __shared__ int s_dsum[kMaxRanges], s_dcnt[kMaxRanges];
  __shared__ Range s_ranges[kMaxRanges];
  const int tile = blockIdx.x;
  const int tid = threadIdx.x;
  int64_t r, c[kWinTPT];
  thread_targets(a.g, tile, tid, r, c);
  if (tid == 0) {
    s_min = INT_MAX;
    s_max = -1;
    s_nr = 0;
    s_fail = 0;
  }
  __syncthreads();
  // pass 1: block span
  int32_t my[kWinTPT][16];
  int lo = INT_MAX, hi = -1;
#pragma unroll
  for (int t = 0; t < kWinTPT; ++t) {
    const bool ok = r < a.g.rows && c[t] < a.g.cols;
    const int64_t j = r * a.g.cols + c[t];
    for (int kk = 0; kk < a.k; ++kk) {
      int32_t v = ok ? a.idx[(int64_t)kk * a.m + j] : -1;
      if (v < 0 || v >= a.n) v = -1;  // sentinel / invalid
      my[t][kk] = v;
      if (v >= 0) {
        lo = min(lo, v / kBlockElems);
        hi = max(hi, v / kBlockElems);
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((tid & 31) == 0) {
    atomicMin(&s_min, lo);
    atomicMax(&s_max, hi);
  }
  __syncthreads();
  const int bmin = s_min, bmax = s_max;
  const int span = (bmax >= 0) ? (bmax - bmin + 1) : 0;
  if (span > kPlanSpanBlocks) {
    if (tid == 0) atomicCAS(&a.status[0], 0, 1);  // indexes too scattered for a window
    return;
  }
  const int words = (span + 31) / 32;
  for (int w = tid; w < words; w += kWinThreads) s_bits[w] = 0u;
  __syncthreads();
#pragma unroll
  for (int t = 0; t < kWinTPT; ++t)
    for (int kk = 0; kk < a.k; ++kk)
      if (my[t][kk] >= 0) {
        const int b = my[t][kk] / kBlockElems - bmin;
        atomicOr(&s_bits[b >> 5], 1u << (b & 31));
      }
  __syncthreads();
  // runs of set blocks -> ranges (sequential walk: a tile has a few dozen runs)
  if (tid == 0) {
    int nr = 0, win = 0, run_start = -1;
    for (int b = 0; b <= span; ++b) {
      const bool set = (b < span) && ((s_bits[b >> 5] >> (b & 31)) & 1u);
      if (set && run_start < 0) run_start = b;
      if (!set && run_start >= 0) {
        if (nr < kMaxRanges) {
          Range rg;
          rg.src_start = (bmin + run_start) * kBlockElems;
          int64_t end = (int64_t)(bmin + b) * kBlockElems;
          if (end > a.n_pad) end = a.n_pad;
          rg.n_elem = (int32_t)(end - rg.src_start);
          rg.smem_off = win;   // provisional: replaced by the skewed offset below
          rg.mask_off = win;   // mask bytes are staged densely (they are only read by the fold pass)
          s_ranges[nr] = rg;
          win += rg.n_elem;
        }
        ++nr;
        run_start = -1;
      }
      // skip whole empty words quickly
      if (run_start < 0 && (b & 31) == 31 && b + 1 < span) {
        while (b + 32 < span && s_bits[(b + 1) >> 5] == 0u) b += 32;
      }
    }
    s_nr = nr;
    s_mwin = win;
    if (nr > kMaxRanges) {
      s_fail = 1;
      atomicCAS(&a.status[0], 0, 2);
    }
  }
  for (int i = tid; i < kMaxRanges; i += kWinThreads) {
    s_dsum[i] = 0;
    s_dcnt[i] = 0;
  }
  __syncthreads();
  if (s_fail) return;
  const int nr = s_nr;
  // ---- range and position of every reference; alignment statistics between consecutive ranges.
  // Consecutive ranges are pieces of neighbouring latitude rings.  When one target references position
  // p_a of range r and p_b of range r+1, the two source points lie at about the same longitude, so
  // D_r = mean(p_a - p_b) tells how the two ranges are shifted against each other.
  int16_t rr[kWinTPT][16];
#pragma unroll
  for (int t = 0; t < kWinTPT; ++t) {
    for (int kk = 0; kk < a.k; ++kk) {
      const int32_t v = my[t][kk];
      int lo_r = -1;
      if (v >= 0) {
        lo_r = 0;
        int hi_r = nr - 1;  // last range with src_start <= v
        while (lo_r < hi_r) {
          const int mid = (lo_r + hi_r + 1) >> 1;
          if (s_ranges[mid].src_start <= v) lo_r = mid; else hi_r = mid - 1;
        }
      }
      rr[t][kk] = (int16_t)lo_r;
    }
    for (int ka = 0; ka < a.k; ++ka)
      for (int kb = 0; kb < a.k; ++kb)
        if (rr[t][ka] >= 0 && rr[t][kb] == rr[t][ka] + 1) {
          const int pa = my[t][ka] - s_ranges[rr[t][ka]].src_start;
          const int pb = my[t][kb] - s_ranges[rr[t][kb]].src_start;
          atomicAdd(&s_dsum[rr[t][ka]], pa - pb);
          atomicAdd(&s_dcnt[rr[t][ka]], 1);
        }
  }
  __syncthreads();
  // ---- skewed shared-memory offsets: with u(r) the position of range r that lines up with position 0 of
  // range 0, a value at "longitude" x = p - u(r) of ring r goes to bank (x + step*(r % mod)) mod 16 (8-byte banks):
  // lanes of a half-warp that read neighbouring rings then hit disjoint bank groups.  (mod, step) = (2, 8) for
  // half-warps that own 16 targets of one row (they see 8-9 points of each of two rings), (4, 4) for 2x8 and 4x4
  // blocks (3-5 points of up to four rings).  O1280 -> EFAS, wavefronts per half-warp gather (tools/bank_sim.py,
  // the first two confirmed by ncu): dense offsets 2.34, 1x16 + (2, 8) 1.62, 2x8 + (4, 4) 1.41, 4x4 + (4, 4) 1.17.
  if (tid == 0) {
    int u = 0, cur = 0;
    for (int i = 0; i < nr; ++i) {
      if (i > 0) {
        const int cnt = s_dcnt[i - 1];
        int d = 0;
        if (cnt > 0) {
          const int sum = s_dsum[i - 1];
          d = (sum >= 0) ? (sum + cnt / 2) / cnt : -((-sum + cnt / 2) / cnt);
        }
        u -= d;
      }
      int want = ((a.skew_step * (i % a.skew_mod) - u) % 16 + 16) % 16;
      want &= ~1;                                   // 16-byte cp.async chunks: even element offsets
      const int off = cur + (((want - cur) % 16) + 16) % 16;
      s_ranges[i].smem_off = off;
      cur = off + s_ranges[i].n_elem;
    }
    const int win = (cur + 15) & ~15;
    if (win > kMaxWindowElems) {
      s_fail = 1;
      atomicCAS(&a.status[0], 0, 3);
    } else {
      atomicMax(&a.status[1], win);
      atomicMax(&a.status[2], nr);
      atomicMax(&a.status[3], s_mwin);
      atomicAdd(a.total_window, (unsigned long long)s_mwin);
      if (!a.dry) {
        a.tile_info[4 * tile] = nr;
        a.tile_info[4 * tile + 1] = win;
        a.tile_info[4 * tile + 2] = s_mwin;
        a.tile_info[4 * tile + 3] = 0;
      }
    }
  }
  __syncthreads();
  if (s_fail) return;
  if (!a.dry)
    for (int i = tid; i < nr; i += kWinThreads) a.ranges[(int64_t)tile * kMaxRanges + i] = s_ranges[i];
  // ---- local offsets; and what the gathers of this layout cost: a 64-bit shared-memory load is served per
  // half-warp, one wavefront per distinct word on the busiest of the 16 8-byte banks (lanes reading the same
  // word share it).  The sum over the tile is what the host compares between half-warp block shapes.
  const int lane = tid & 31, hw_base = lane & 16, hl = lane & 15;
  unsigned wave_sum = 0;
#pragma unroll
  for (int t = 0; t < kWinTPT; ++t) {
    const bool ok = r < a.g.rows && c[t] < a.g.cols;
    const int64_t j = r * a.g.cols + c[t];
    uint16_t lo[16];
    for (int kk = 0; kk < a.k; ++kk) {
      uint16_t l = kLocalSentinel;   // the sentinel slot is the last element of a stage: bank 15, like 0xffff
      if (rr[t][kk] >= 0) l = (uint16_t)(8 * (s_ranges[rr[t][kk]].smem_off + (my[t][kk] - s_ranges[rr[t][kk]].src_start)));
      lo[kk] = l;
    }
    // ---- the order in which a target loads neighbours 0 / 1 and 2 / 3 is free: the sequential sum starts with
    // w0*v0 + w1*v1 and IEEE addition commutes, so the kernel may exchange the two (weights with them) without
    // changing a bit of the result; for 2 / 3 the kernel loads them exchanged and puts the two products back in
    // distance order with a select before they are added.  pick_pair_order() chooses per half-warp.  Bit 0 of the
    // pair's first offset (byte offsets are multiples of 8) tells the kernel that the pair is exchanged.
    // O1280 -> EFAS, 2x8 blocks, wavefronts of the four gathers: 1.11 1.38 1.47 1.78 (mean 1.43) -> 1.01 1.06 1.07 1.12.
    if (a.k >= 2 && a.swap01) {
      for (int pr = 0; pr < ((a.k == 4 && a.swap01 >= 2) ? 2 : 1); ++pr) {
        unsigned A = lo[2 * pr] >> 3, B = lo[2 * pr + 1] >> 3;
        const bool can = lo[2 * pr] != kLocalSentinel && lo[2 * pr + 1] != kLocalSentinel;
        if (pick_pair_order(A, B, can, hw_base, hl, a.swap_sweeps)) {
          const uint16_t tmp = lo[2 * pr];
          lo[2 * pr] = lo[2 * pr + 1] | 1u;
          lo[2 * pr + 1] = tmp;
        }
      }
    }
    for (int kk = 0; kk < a.k; ++kk) {
      const uint16_t l = lo[kk];
      if (ok && !a.dry) a.lidx[(int64_t)kk * a.m + j] = l;
      const unsigned mine = l >> 3;  // the 8-byte word
      bool first = true;  // no lower lane of the half-warp reads the same word
      for (int q = 0; q < 16; ++q) {
        const unsigned other = __shfl_sync(0xffffffffu, mine, hw_base + q);
        if (q < hl && other == mine) first = false;
      }
      int cnt = 0;        // distinct words on this lane's bank
      for (int q = 0; q < 16; ++q) {
        const unsigned other = __shfl_sync(0xffffffffu, mine, hw_base + q);
        const int f = __shfl_sync(0xffffffffu, (int)first, hw_base + q);
        if (f && ((other ^ mine) & 15u) == 0) ++cnt;
      }
      for (int o = 8; o > 0; o >>= 1) cnt = max(cnt, __shfl_xor_sync(0xffffffffu, cnt, o));
      wave_sum += (unsigned)cnt;
    }
  }
  if (hl == 0) atomicAdd(a.total_wavefronts, (unsigned long long)wave_sum);
}

// ------------------------------------------------------------------ PTX helpers: cp.async (LDGSTS)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 16-byte global -> shared asynchronous copy, L2 only (.cg): the staged values are read once per tile
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src_gmem) : "memory");
}
// ... with an L2 eviction-priority hint: the staged source values are read again by the neighbouring tiles (windows
// overlap 2.4x on O1280 -> EFAS) while the outputs stream through the same L2
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void cp_async16_hint(uint32_t dst_smem, const void* src_gmem, unsigned long long pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src_gmem), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// mbarrier (shared-memory barrier objects) for the lagged message loop of the FAST marked-field kernels
__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
// the arrival is triggered when every cp.async this thread has issued so far has landed (counts as one expected arrival)
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t addr) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "G2P_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra G2P_MBAR_DONE;\n"
      "bra G2P_MBAR_WAIT;\n"
      "G2P_MBAR_DONE:\n"
      "}" ::"r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
// the same with the window offset taken from the low / high half of a packed pair; the unpacking sits inside the
// volatile asm so that it stays in the message loop (hoisted, it would cost the registers the packing saves)
__device__ __forceinline__ double lds_f64_lo16(uint32_t base, uint32_t packed) {
  double v;
  asm volatile("{\n\t.reg .u32 t;\n\tand.b32 t, %2, 0xffff;\n\tadd.u32 t, t, %1;\n\tld.shared.f64 %0, [t];\n\t}" : "=d"(v) : "r"(base), "r"(packed));
  return v;
}
__device__ __forceinline__ double lds_f64_hi16(uint32_t base, uint32_t packed) {
  double v;
  asm volatile("{\n\t.reg .u32 t;\n\tshr.u32 t, %2, 16;\n\tadd.u32 t, t, %1;\n\tld.shared.f64 %0, [t];\n\t}" : "=d"(v) : "r"(base), "r"(packed));
  return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// one output message of the fused pipeline: which input messages it combines and how (g2p_manip_item, <= 2 inputs)
struct AggItem {
  int32_t kind, in0, in1, pad;
  double unit_time, divisor;
};

// ------------------------------------------------------------------ apply kernel
struct WinArgs {
  const uint16_t* lidx;
  const double* coef;
  const Range* ranges;
  const int32_t* tile_info;
  const double* src;
  const uint8_t* src_mask;
  double* out;
  uint8_t* out_mask;
  int64_t m, src_stride, out_stride;
  int64_t sbits_stride;   // MP == 3: bytes per source mask row (src_stride / 8)
  int64_t obits_stride;   // MP >= 3: uint32 words per output mask row
  int pairwise;           // G2P_SUM_PAIRWISE
  int obits_ballot;       // MP >= 3: mask bytes assembled from warp ballots (else atomicOr into cleared rows)
  TileGeom g;
  int b, n_tiles, msgs_per_chunk, stages, stage_elems, stage_melems;
  int64_t n_items;
  double mv_out;
  const double* corr_gem;
  const double* corr_dem_term;
  const uint8_t* corr_valid;
  double corr_mv;
  int corr_mode;  // g2p_correction.mode
  // writers' post-processing (g2p_writer_post), after the correction
  int post_on, post_use_match;
  int post_f32;           // CORR: the fields leave as float32 rows of out_stride values (g2p_writer_post.out_f32)
  const uint8_t* post_clone;
  double post_mv_match, post_mv_write;
  // AGG: the source-grid stages fused into the staging step (see apply_window_kernel)
  const AggItem* agg_items;   // device, one per output message
  int op1, op2, cut_off;
  double c1, c2, mv_grib;
};

// pairwise (block-uniform, K = 4): (w0*v0 + w2*v2) + (w1*v1 + w3*v3), see G2P_SUM_PAIRWISE
// x23 (K = 4, sequential order): neighbours 2 and 3 were loaded exchanged (weights too): their products go back in
// distance order before they are added
template <int K>
__device__ __forceinline__ double weighted_sum(const double (&w)[K], const double (&v)[K], int pairwise, bool x23) {
  if (K == 1) return v[0];
  if (K == 4 && pairwise)
    return __dadd_rn(__dadd_rn(__dmul_rn(w[0], v[0]), __dmul_rn(w[K > 2 ? 2 : 0], v[K > 2 ? 2 : 0])),
                     __dadd_rn(__dmul_rn(w[1], v[1]), __dmul_rn(w[K > 3 ? 3 : 0], v[K > 3 ? 3 : 0])));
  if (K == 4) {
    const double q2 = __dmul_rn(w[K > 2 ? 2 : 0], v[K > 2 ? 2 : 0]), q3 = __dmul_rn(w[K > 3 ? 3 : 0], v[K > 3 ? 3 : 0]);
    const double acc = __dadd_rn(__dmul_rn(w[0], v[0]), __dmul_rn(w[1], v[1]));
    return __dadd_rn(__dadd_rn(acc, x23 ? q3 : q2), x23 ? q2 : q3);
  }
  double acc = __dmul_rn(w[0], v[0]);
#pragma unroll
  for (int kk = 1; kk < K; ++kk) acc = __dadd_rn(acc, __dmul_rn(w[kk], v[kk]));
  return acc;
}

// source element staged at value-window element `elem` (ranges are sorted by smem_off); -1 in a gap
__device__ __forceinline__ int value_window_to_source(const Range* ranges, int nr, int elem) {
  int lo = 0, hi = nr - 1;  // last range with smem_off <= elem
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (ranges[mid].smem_off <= elem) lo = mid; else hi = mid - 1;
  }
  const int d = elem - ranges[lo].smem_off;
  return (nr > 0 && d >= 0 && d < ranges[lo].n_elem) ? ranges[lo].src_start + d : -1;
}
// range that holds mask-window element `elem` (dense offsets); -1 past the window
__device__ __forceinline__ int mask_window_range(const Range* ranges, int nr, int mwindow, int elem) {
  if (elem >= mwindow) return -1;
  int lo = 0, hi = nr - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (ranges[mid].mask_off <= elem) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// A masked source value is replaced, in shared memory, by this quiet NaN: the weighted sum of a target
// with a masked neighbour is then NaN and the (rare) slow path looks for the marker among its operands,
// so the k mask-byte gathers per target disappear from the inner loop.
constexpr unsigned long long kMaskMarker = G2P_MASK_MARKER;
// The test for "masked" looks at the upper word only: one integer compare per gathered operand, no branch (with
// scattered missing points nearly every warp holds a masked target, so a "rare path" is not rare).  The fold passes
// only write the upper word over a masked value.  (Deriving the mask from the NaN payload of the weighted sum does
// not work: sm_100a's DADD does not reliably pass the payload on - measured, 57 % of the warps fell through.)
constexpr uint32_t kMarkerHi = (uint32_t)(G2P_MASK_MARKER >> 32);
__device__ __forceinline__ void sts_marker(uint32_t value_addr) {
  asm volatile("st.shared.u32 [%0+4], %1;" ::"r"(value_addr), "r"(kMarkerHi) : "memory");
}

// MP: 0 = no masks (AS_REFERENCE), 1 = PROPAGATE (byte masks in and out), 2 = FILL (byte masks in),
//     3 = PROPAGATE_BITS: the source mask is one bit per point.  Nothing of it is staged: the ranges are 16-element
//         blocks, so every block's mask is one aligned uint16 of the message's bit row; the thread that owns a block
//         loads it into a register one message ahead (a plain read-only load, its latency hidden behind a whole
//         message of gathers) and the fold pass writes markers for its set bits.  The output mask is bit-packed
//         too: the launcher clears it and only MASKED outputs touch it (atomicOr in the rare NaN path), so the hot
//         loop has no mask stores at all.
//     4 = MARKED: the source values already carry the marker: no mask input, no fold pass; out_mask as for 3
//         (or none).
// CV: 16-byte value chunks copied per thread and message (value window <= 512*CV elements)
// S : stages
// where(x != mv, op2(op1(x)), mv)  (manipulation/conversion.py:21); op1 == NONE is the identity 'x=x'
__device__ __forceinline__ double agg_convert(const WinArgs& a, double x) {
  if (a.op1 == G2P_OP_NONE) return x;
  if (x == a.mv_grib) return a.mv_grib;
  double y = a.op1 == G2P_OP_MUL ? __dmul_rn(x, a.c1) : a.op1 == G2P_OP_ADD ? __dadd_rn(x, a.c1) : __ddiv_rn(x, a.c1);
  if (a.op2 != G2P_OP_NONE)
    y = a.op2 == G2P_OP_MUL ? __dmul_rn(y, a.c2) : a.op2 == G2P_OP_ADD ? __dadd_rn(y, a.c2) : __ddiv_rn(y, a.c2);
  return y;
}

// CORR: epilogue variant: the Corrector formula `where(valid & (p != mv), (p + gem) - dem_term, mv)`
//       (manipulation/correction.py:46,74-81) when corr_gem is set, then the writers' post-processing
//       (clone mask | value mask | NaN | == mv_output -> the file's missing value) when post_on is set
// AGG : the source-grid stages run inside the staging step: an output message stages the windows of its (<= 2)
//       input messages, and the fold pass - one message ahead of the gathers, where the mask bytes are folded
//       anyway - evaluates conversion, aggregation (PICK / ACCUM / AVERAGE of <= 2 / COPY, g2p_manip's arithmetic)
//       and the cut-off on every staged source value before the targets gather it.  convert -> aggregate ->
//       cut-off -> interpolate -> correct is then ONE launch that reads every needed source value once.
// FAST: the block-uniform run-time switches of the message loop resolved at compile time for the resident-field paths
//       (MP >= 3, no epilogue): 1 = sequential sum, bit mask rows from ballots; 2 = sequential sum, no mask rows;
//       0 = everything read from WinArgs (pairwise sums, atomicOr mask rows, and every other variant).  The loop of
//       the headline launch drops from 210 to ~165 instructions per thread and message (profiles/r02_apply.md).
template <int K, int MP, int CV, int S, bool CORR, bool AGG, int FAST = 0>
__global__ void __launch_bounds__(kWinThreads, (CORR || AGG) ? 2 : 3) apply_window_kernel(const WinArgs a) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ Range s_ranges[kMaxRanges];
  __shared__ int s_tile_info[4];
  constexpr bool MASK = MP == 1 || MP == 2;  // byte masks staged through shared memory
  constexpr bool BITS = MP == 3;             // bit masks read into registers
  constexpr bool ANYMASK = MP != 0;          // a result can be masked
  constexpr bool OBITS = MP >= 3;            // bit-packed output mask
  constexpr bool FOLD = MASK || BITS || AGG;
  constexpr int NIN = AGG ? 2 : 1;  // staged input windows per message
  static_assert(FAST == 0 || (MP >= 3 && !CORR && !AGG), "FAST variants exist for the resident-field paths only");
  const int pairwise = FAST ? 0 : a.pairwise;
  // LAG: the per-message __syncthreads replaced by two rings of mbarriers with one message of slack.  full[s]: the
  // copies of the message in stage s have landed (every thread's cp.async arrival); empty[s]: every warp is done
  // gathering from stage s.  A warp that is ahead runs into the next message instead of waiting at a CTA barrier for
  // the slowest one.  Copies run S-2 = 3 messages ahead and land in the stage of message it-2.  Measured on
  // O1280 -> EFAS (profiles/r02_apply_lag.md): 3 messages ahead is the optimum in both schemes (2 and 4 lose 3-8 %),
  // rings longer than look-ahead + 2 and stages cut to each tile's own window gain nothing.
  constexpr bool LAG = FAST != 0 && MP == 4 && S == 5;
  __shared__ __align__(8) unsigned long long s_mbar[LAG ? 2 * S : 2];
  const uint32_t mb_full = smem_u32(s_mbar), mb_empty = mb_full + 8u * S;
  // stage cursors and phase parities of the copies (p) and the gathers (g), kept across items: every item leaves the
  // barriers quiescent (as many messages gathered as copied)
  uint32_t lag_ps = 0, lag_pp = 0, lag_gs = 0, lag_gp = 0;
  const unsigned long long l2_keep = LAG ? l2_policy_evict_last() : 0ull;
  if (LAG && threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < S; ++i) {
      mbar_init(mb_full + 8u * i, kWinThreads);
      mbar_init(mb_empty + 8u * i, kWinThreads / 32);
    }
  }
  constexpr int CM = (CV + 7) / 8;  // 16-byte mask chunks per thread; also 4-element fold groups per thread / 4
  constexpr int CF = (CV + 1) / 2;  // fold groups (4 mask bytes) per thread
  constexpr int CB = (CV + 3) / 4;  // 8-element groups (one byte of mask bits) per thread
  static_assert(!(AGG && MP >= 3), "the fused kernel takes byte masks");
  const int tid = threadIdx.x;
  const uint32_t val_base = smem_u32(s_raw);                                             // [S][NIN][stage_elems] f64
  const uint32_t msk_base = val_base + (uint32_t)S * NIN * a.stage_elems * 8u;           // [S][NIN][stage_melems] u8
  const uint32_t in_vbytes = (uint32_t)a.stage_elems * 8u, in_mbytes = (uint32_t)a.stage_melems;
  const uint32_t stage_vbytes = NIN * in_vbytes;
  const uint32_t stage_mbytes = NIN * in_mbytes;
  const uint32_t val_end = val_base + (uint32_t)S * stage_vbytes;
  const int sent = a.stage_elems - 1;  // slot that always holds mv_out (copies and folds never reach it)
  if (tid < S) reinterpret_cast<double*>(s_raw)[(size_t)tid * NIN * a.stage_elems + sent] = a.mv_out;
  const int64_t vrow = a.src_stride * 8, mrow = a.src_stride;

  for (int64_t item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    const int tile = (int)(item % a.n_tiles);
    const int chunk = (int)(item / a.n_tiles);
    const int msg_begin = chunk * a.msgs_per_chunk;
    const int n_msg = min(a.b, msg_begin + a.msgs_per_chunk) - msg_begin;
    // ---- tile plan
    __syncthreads();  // previous item is finished with s_ranges and the stages
    if (tid < 4) s_tile_info[tid] = a.tile_info[4 * tile + tid];
    __syncthreads();
    const int nr = s_tile_info[0];
    const int mwindow = s_tile_info[2];
    for (int i = tid; i < nr; i += kWinThreads) s_ranges[i] = a.ranges[(int64_t)tile * kMaxRanges + i];
    __syncthreads();
    // ---- this thread's share of the staging work, as byte offsets (-1 = nothing):
    //   vsrc[i]: source-row offset of the 16-byte value chunk that goes to stage bytes 16*(tid + 256 i)
    //   msrc[i]: source-row offset of the 16-byte mask chunk that goes to mask-stage bytes 16*(tid + 256 i)
    //   fdst[i]: stage byte offset of the 4 values whose mask bytes are mask-stage bytes 4*(tid + 256 i)
    int vsrc[CV], msrc[CM], fdst[CF];
#pragma unroll
    for (int i = 0; i < CV; ++i) {
      const int e = value_window_to_source(s_ranges, nr, 2 * (tid + kWinThreads * i));
      vsrc[i] = e < 0 ? -1 : 8 * e;
    }
#pragma unroll
    for (int i = 0; i < CM; ++i) {
      msrc[i] = -1;
      if (MASK) {
        const int e = 16 * (tid + kWinThreads * i);
        const int rg = mask_window_range(s_ranges, nr, mwindow, e);
        if (rg >= 0) msrc[i] = s_ranges[rg].src_start + (e - s_ranges[rg].mask_off);
      }
    }
#pragma unroll
    for (int i = 0; i < CF; ++i) {
      fdst[i] = -1;
      if (MASK || AGG) {
        const int e = 4 * (tid + kWinThreads * i);
        const int rg = mask_window_range(s_ranges, nr, mwindow, e);
        if (rg >= 0) fdst[i] = 8 * (s_ranges[rg].smem_off + (e - s_ranges[rg].mask_off));
      }
    }
    //   bsrc[i]: byte index, in a message's bit row, of the mask bits of the 8 values staged at stage bytes bdst[i]
    //   (the ranges are 16-element blocks, so every aligned group of 8 staged values is one byte of the bit row; the
    //   work is spread like the byte-mask fold: one group per thread, no loops)
    int bsrc[BITS ? CB : 1], bdst[BITS ? CB : 1];
    if (BITS) {
#pragma unroll
      for (int i = 0; i < CB; ++i) {
        bsrc[i] = -1;
        bdst[i] = 0;
        const int e = 8 * (tid + kWinThreads * i);
        const int rg = mask_window_range(s_ranges, nr, mwindow, e);
        if (rg >= 0) {
          bsrc[i] = (s_ranges[rg].src_start + (e - s_ranges[rg].mask_off)) >> 3;
          bdst[i] = 8 * (s_ranges[rg].smem_off + (e - s_ranges[rg].mask_off));
        }
      }
    }
    const uint8_t* brow = BITS ? a.src_mask + (int64_t)msg_begin * a.sbits_stride : nullptr;
    uint32_t bcur[BITS ? CB : 1], bnxt[BITS ? CB : 1];
    auto load_bits = [&](uint32_t (&dst)[BITS ? CB : 1], int msg) {   // message `msg` of the chunk (zeros past its end)
#pragma unroll
      for (int i = 0; i < (BITS ? CB : 1); ++i)
        dst[i] = (BITS && bsrc[i] >= 0 && msg < n_msg) ? (uint32_t)__ldg(brow + (int64_t)msg * a.sbits_stride + bsrc[i]) : 0u;
    };
    if (BITS) {
      load_bits(bcur, 0);
      load_bits(bnxt, 1);
    }
    const char* grow_v = reinterpret_cast<const char*>(a.src) + (int64_t)msg_begin * vrow;        // block-uniform
    const char* grow_m = MASK ? reinterpret_cast<const char*>(a.src_mask) + (int64_t)msg_begin * mrow : nullptr;
    int n_left = n_msg;  // messages not yet prefetched

    // next message of the chunk -> the stage whose value / mask addresses are (sv, sm)
    auto prefetch = [&](uint32_t sv, uint32_t sm) {
      if (n_left > 0) {
        if (AGG) {
          const AggItem it = a.agg_items[msg_begin + (n_msg - n_left)];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int in = q == 0 ? it.in0 : it.in1;
            if (in < 0) continue;
            const char* gv = reinterpret_cast<const char*>(a.src) + (int64_t)in * vrow;
#pragma unroll
            for (int i = 0; i < CV; ++i)
              if (vsrc[i] >= 0) cp_async16(sv + q * in_vbytes + 16u * tid + 16u * kWinThreads * i, gv + vsrc[i]);
            if (MASK) {
              const char* gm = reinterpret_cast<const char*>(a.src_mask) + (int64_t)in * mrow;
#pragma unroll
              for (int i = 0; i < CM; ++i)
                if (msrc[i] >= 0) cp_async16(sm + q * in_mbytes + 16u * tid + 16u * kWinThreads * i, gm + msrc[i]);
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < CV; ++i)
            if (vsrc[i] >= 0) cp_async16(sv + 16u * tid + 16u * kWinThreads * i, grow_v + vsrc[i]);
          if (MASK) {
#pragma unroll
            for (int i = 0; i < CM; ++i)
              if (msrc[i] >= 0) cp_async16(sm + 16u * tid + 16u * kWinThreads * i, grow_m + msrc[i]);
            grow_m += mrow;
          }
          grow_v += vrow;
        }
        --n_left;
      }
      cp_async_commit();  // always: keeps the group count uniform
    };
    // the pass over a landed message, one message ahead of the gathers: (AGG) combine the staged input windows
    // into the values the targets will gather, then fold the mask bytes into them
    int fold_it = 0;  // message of the chunk the next fold works on
    auto fold = [&](uint32_t sv, uint32_t sm) {
      if (BITS) {
#pragma unroll
        for (int i = 0; i < CB; ++i) {
          const uint32_t bits = bcur[i];
          if (bits) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if ((bits >> q) & 1u)
                sts_marker(sv + (uint32_t)bdst[i] + 8u * q);
          }
        }
        return;
      }
      AggItem it{};
      if (AGG) it = a.agg_items[msg_begin + fold_it];
      ++fold_it;
#pragma unroll
      for (int i = 0; i < CF; ++i) {
        if (fdst[i] < 0) continue;
        if (AGG) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t at = sv + (uint32_t)fdst[i] + 8u * q;
            const double xa = it.in0 >= 0 ? agg_convert(a, lds_f64(at)) : 0.0;
            const double xb = it.in1 >= 0 ? agg_convert(a, lds_f64(at + in_vbytes)) : 0.0;
            double v;
            if (it.kind == G2P_AGG_ACCUM) {
              v = __ddiv_rn(__dmul_rn(__dsub_rn(__dadd_rn(0.0, xa), xb), it.unit_time), it.divisor);
            } else if (it.kind == G2P_AGG_AVERAGE) {
              v = __dadd_rn(0.0, xa);
              if (it.in1 >= 0) v = __dadd_rn(v, xb);
              v = __ddiv_rn(v, it.divisor);
            } else if (it.kind == G2P_AGG_COPY) {
              v = xa;
            } else {
              v = it.in0 >= 0 ? __dadd_rn(0.0, xa) : 0.0;
            }
            if (a.cut_off && v < 0.0) v = 0.0;
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(at), "d"(v) : "memory");
          }
        }
        if (MASK) {
          uint32_t m4 = 0;
          const uint32_t ma = sm + 4u * tid + 4u * kWinThreads * i;
          if (!AGG) {
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(m4) : "r"(ma));
          } else {
            // ACCUM keeps only the mask of V[t-D], AVERAGE the OR of its inputs, COPY its input's, PICK none
            uint32_t m_a = 0, m_b = 0;
            if (it.in0 >= 0) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(m_a) : "r"(ma));
            if (it.in1 >= 0) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(m_b) : "r"(ma + in_mbytes));
            m4 = it.kind == G2P_AGG_ACCUM ? m_b : it.kind == G2P_AGG_AVERAGE ? (m_a | m_b) : it.kind == G2P_AGG_COPY ? m_a : 0u;
          }
          if (m4) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if ((m4 >> (8 * q)) & 0xffu)
                sts_marker(sv + (uint32_t)fdst[i] + 8u * q);
          }
        }
      }
    };
    // running stage cursors (value address, mask address)
    uint32_t pv = val_base, pm = msk_base;  // stage the next prefetch goes to
    auto advance = [&](uint32_t& v, uint32_t& mk) {
      v += stage_vbytes;
      mk += stage_mbytes;
      if (v == val_end) {
        v = val_base;
        mk = msk_base;
      }
    };
    // LAG: next message of the chunk -> stage lag_ps, once every warp has left the message that used it
    auto prefetch_lag = [&]() {
      if (n_left > 0) {
        mbar_wait(mb_empty + 8u * lag_ps, lag_pp ^ 1u);
        const uint32_t sv = val_base + lag_ps * stage_vbytes;
#pragma unroll
        for (int i = 0; i < CV; ++i)
          if (vsrc[i] >= 0) cp_async16_hint(sv + 16u * tid + 16u * kWinThreads * i, grow_v + vsrc[i], l2_keep);
        cp_async_mbar_arrive(mb_full + 8u * lag_ps);
        grow_v += vrow;
        --n_left;
        if (++lag_ps == S) {
          lag_ps = 0;
          lag_pp ^= 1u;
        }
      }
    };
    if (LAG) {
#pragma unroll
      for (int s = 0; s < S - 2; ++s) prefetch_lag();
    } else {
#pragma unroll
      for (int s = 0; s < S - 1; ++s) {
        prefetch(pv, pm);
        advance(pv, pm);
      }
    }

    // ---- this thread's targets: window offsets and weights stay in registers for the whole chunk
    int64_t r, c[kWinTPT];
    thread_targets(a.g, tile, tid, r, c);
    bool ok[kWinTPT];
    uint32_t x23 = 0;   // bit t: slot t loads neighbours 2 and 3 exchanged
    // The epilogue variants keep neighbours 2 / 3 in distance order (the offsets are put back at the item load): with
    // the selects in their loop the configs[4] pass went from 0.448 to 0.480 ms (bench `pipeline`, also with a plan
    // that exchanges nothing), so the loop of these kernels is the one measured at 0.448.  Parity of this state:
    // the epilogue / fused tests on the GPU; its pass time has not been measured again (GPU budget of the round spent).
    constexpr bool X23 = !(CORR || AGG);
    // window byte offsets of the neighbours, two uint16 per register (K = 4): the kernel lives at the 80-register
    // line of three CTAs per SM, and eight registers saved here are what the bit-mask paths need
    constexpr int KP = K > 1 ? K / 2 : 1;
    uint32_t li[kWinTPT][KP];
    double w[kWinTPT][K];
    const int64_t j0 = r * a.g.cols + c[0];
    uint32_t ob[kWinTPT];  // byte offset of the target in an output row
    double cgem[CORR ? kWinTPT : 1], cdt[CORR ? kWinTPT : 1];
    bool cvalid[CORR ? kWinTPT : 1], pclone[CORR ? kWinTPT : 1];
#pragma unroll
    for (int t = 0; t < kWinTPT; ++t) {
      ok[t] = r < a.g.rows && c[t] < a.g.cols;
      const int64_t j = j0 + (int64_t)t * a.g.threads_per_row;
      ob[t] = 8u * (uint32_t)j;
      uint16_t lr[K];
#pragma unroll
      for (int kk = 0; kk < K; ++kk) {
        lr[kk] = ok[t] ? __ldg(a.lidx + (int64_t)kk * a.m + j) : kLocalSentinel;
        w[t][kk] = (K > 1 && ok[t]) ? __ldg(a.coef + (int64_t)kk * a.m + j) : 1.0;
      }
      if (K == 4 && lr[K > 2 ? 2 : 0] != kLocalSentinel && (lr[K > 2 ? 2 : 0] & 1u)) {
        // ... and neighbours 2 / 3 (bit 0 of the third offset): loaded exchanged, with the weights; weighted_sum puts the
        // two products back in distance order
        constexpr int i2 = K > 2 ? 2 : 0, i3 = K > 3 ? 3 : 0;
        lr[i2] &= (uint16_t)~1u;
        if (pairwise || !X23) {   // distance order again
          const uint16_t tmp = lr[i2];
          lr[i2] = lr[i3];
          lr[i3] = tmp;
        } else {
          const double tmp = w[t][i2];
          w[t][i2] = w[t][i3];
          w[t][i3] = tmp;
          x23 |= 1u << t;
        }
      }
      if (K >= 2 && lr[0] != kLocalSentinel && (lr[0] & 1u)) {
        // the plan exchanged the two nearest neighbours (bit 0 of the first offset): w0*v0 + w1*v1 commutes, so the
        // sequential sum takes the weights exchanged too; the pairwise order puts the offsets back
        lr[0] &= (uint16_t)~1u;
        if (pairwise) {
          const uint16_t tmp = lr[0];
          lr[0] = lr[K > 1 ? 1 : 0];
          lr[K > 1 ? 1 : 0] = tmp;
        } else {
          const double tmp = w[t][0];
          w[t][0] = w[t][K > 1 ? 1 : 0];
          w[t][K > 1 ? 1 : 0] = tmp;
        }
      }
#pragma unroll
      for (int kk = 0; kk < K; ++kk) {
        const uint32_t lb = (lr[kk] == kLocalSentinel) ? 8u * (uint32_t)sent : (uint32_t)lr[kk];  // byte offsets as the plan stores them
        if (K == 1) li[t][0] = lb;
        else if ((kk & 1) == 0) li[t][kk >> 1] = lb;
        else li[t][kk >> 1] |= lb << 16;
      }
      if (CORR) {
        const bool cr = ok[t] && a.corr_gem != nullptr;
        cgem[t] = cr ? __ldg(a.corr_gem + j) : 0.0;
        cdt[t] = cr ? __ldg(a.corr_dem_term + j) : 0.0;
        cvalid[t] = cr ? __ldg(a.corr_valid + j) != 0 : false;
        pclone[t] = (ok[t] && a.post_clone) ? __ldg(a.post_clone + j) != 0 : false;
      }
    }
    // bytes per output value: 8, or 4 when the writers' epilogue hands over float32 fields (CORR variants only)
    const int out_vbytes = (CORR && a.post_f32) ? 4 : 8;
    char* o = reinterpret_cast<char*>(a.out) + (int64_t)msg_begin * a.out_stride * out_vbytes;       // block-uniform
    char* om = (MP == 1) ? reinterpret_cast<char*>(a.out_mask) + (int64_t)msg_begin * a.out_stride : nullptr;
    uint32_t* ob_row = (OBITS && FAST != 2 && (FAST == 1 || a.out_mask)) ? reinterpret_cast<uint32_t*>(a.out_mask) + (int64_t)msg_begin * a.obits_stride : nullptr;
    const bool use_ballot = FAST == 1 ? true : FAST == 2 ? false : (OBITS && ob_row != nullptr && a.obits_ballot);
    const bool use_atomic = FAST != 0 ? false : (OBITS && ob_row != nullptr && !a.obits_ballot);
    // Bit-packed output mask from ballots: per target slot a warp owns whole bytes of the mask rows (8 neighbouring
    // columns of one row sit in 8 neighbouring lanes - two nibbles for 4x4 blocks), 4 bytes per slot, 16 per
    // message.  Lane q < 16 writes byte (slot q >> 2, byte q & 3): its row offset and the position of its bits in
    // the slot's ballot are fixed for the whole chunk.
    if (use_ballot && tile == a.n_tiles - 1) {   // the padding bytes of the chunk's mask rows
      const int used = (int)((a.m + 7) >> 3), tail = (int)(a.obits_stride * 4) - used;
      for (int i = tid; i < tail * n_msg; i += kWinThreads)
        reinterpret_cast<uint8_t*>(ob_row)[(int64_t)(i / tail) * a.obits_stride * 4 + used + (i % tail)] = 0;
    }
    int ob_byte = -1;        // byte offset in a mask row (-1: nothing to write)
    uint32_t ob_shift = 0;   // lane of the byte's first bit
    if (use_ballot) {
      const int lane = tid & 31;
      const int bc_small = a.g.block_cols_log2 == 2;
#pragma unroll
      for (int t = 0; t < kWinTPT; ++t) {
        const int mine = ok[t] ? (int)(ob[t] >> 6) : -1;   // byte of this lane's target
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int leader = bc_small ? 4 * i : 8 * i;
          const int v = __shfl_sync(0xffffffffu, mine, leader);
          if (lane == 4 * t + i) {
            ob_byte = v;
            ob_shift = (uint32_t)leader;
          }
        }
      }
    }

    uint32_t gv_s = LAG ? val_base + lag_gs * stage_vbytes : val_base;   // stage of the message being gathered
    uint32_t fv_s = val_base, fm_s = msk_base;  // stage of the message being folded (one ahead)
    if (FOLD) {  // message 0 must be folded before the first gather
      cp_async_wait<S - 2>();
      __syncthreads();
      fold(fv_s, fm_s);
      advance(fv_s, fm_s);
      if (BITS) {
#pragma unroll
        for (int i = 0; i < CB; ++i) bcur[i] = bnxt[i];  // bits of message 1
        load_bits(bnxt, 2);
      }
    }
    // ---- messages
    for (int it = 0; it < n_msg; ++it) {
      // masked: message it+1 has landed (it is folded below, one message ahead of the gathers); else message it
      if (LAG) {
        prefetch_lag();       // message it+S-2, into the stage message it-2 used
        mbar_wait(mb_full + 8u * lag_gs, lag_gp);   // everyone's copies of message it have landed
      } else {
        cp_async_wait<FOLD ? S - 3 : S - 2>();
        __syncthreads();        // copies / folds of the others are visible; everyone is done with message it-1
        prefetch(pv, pm);       // message it+S-1, into the stage message it-1 used
        advance(pv, pm);
      }
      if (FOLD) {
        if (it + 1 < n_msg) fold(fv_s, fm_s);
        advance(fv_s, fm_s);
      }
      if (BITS) {
        // rotate: the load issued one message ago is consumed here, the next one has a whole message to land
#pragma unroll
        for (int i = 0; i < CB; ++i) bcur[i] = bnxt[i];
        load_bits(bnxt, it + 3);
      }
      uint32_t mkbits = 0;   // OBITS: bit t = slot t's output is masked
#pragma unroll
      for (int t = 0; t < kWinTPT; ++t) {
        double v[K];
#pragma unroll
        for (int kk = 0; kk < K; ++kk)
          v[kk] = K == 1 ? lds_f64(gv_s + li[t][0]) : (kk & 1) ? lds_f64_hi16(gv_s, li[t][kk >> 1]) : lds_f64_lo16(gv_s, li[t][kk >> 1]);
        double res = weighted_sum<K>(w[t], v, pairwise, X23 && ((x23 >> t) & 1u));
        uint32_t mk = 0;
        if (ANYMASK) {
          // branch-free: with scattered missing points nearly every warp holds a masked target
#pragma unroll
          for (int kk = 0; kk < K; ++kk) mk |= (uint32_t)((uint32_t)__double2hiint(v[kk]) == kMarkerHi);
          if (mk) res = a.mv_out;
          if (use_atomic && mk && ok[t]) atomicOr(ob_row + (ob[t] >> 8), 1u << ((ob[t] >> 3) & 31u));
        }
        if (OBITS && FAST != 2) mkbits |= (mk != 0 && ok[t]) ? (1u << t) : 0u;
        if (CORR) {
          if (a.corr_gem && !mk)  // a masked output stays the missing value; p = NaN passes the guards and stays NaN
            res = !(cvalid[t] && res != a.corr_mv) ? a.corr_mv
                  : a.corr_mode == 0 ? __dsub_rn(__dadd_rn(res, cgem[t]), cdt[t]) : __dmul_rn(__ddiv_rn(res, cgem[t]), cdt[t]);
          if (a.post_on && (pclone[t] || mk || res != res || (a.post_use_match && res == a.post_mv_match))) res = a.post_mv_write;
        }
        if (ok[t]) {
          if (CORR && a.post_f32)
            __stcs(reinterpret_cast<float*>(o + (ob[t] >> 1)), (float)res);   // round to nearest, as a C / numpy cast
          else
            __stcs(reinterpret_cast<double*>(o + ob[t]), res);
          if (MP == 1) *reinterpret_cast<uint8_t*>(om + (ob[t] >> 3)) = (uint8_t)mk;
        }
      }
      if (use_ballot) {   // block-uniform; the votes come after the four slots' arithmetic
        const int slot = (tid & 31) >> 2;
        uint32_t bt = 0;
#pragma unroll
        for (int t = 0; t < kWinTPT; ++t) {
          const uint32_t v = __ballot_sync(0xffffffffu, (mkbits >> t) & 1u);
          if (slot == t) bt = v;
        }
        const uint32_t byte = a.g.block_cols_log2 == 2 ? (((bt >> ob_shift) & 0xfu) | (((bt >> (ob_shift + 16)) & 0xfu) << 4))
                                                       : ((bt >> ob_shift) & 0xffu);
        if (ob_byte >= 0) reinterpret_cast<uint8_t*>(ob_row)[ob_byte] = (uint8_t)byte;
      }
      if (LAG) {   // this warp is done with the stage
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(mb_empty + 8u * lag_gs);
        if (++lag_gs == S) {
          lag_gs = 0;
          lag_gp ^= 1u;
        }
      }
      gv_s += stage_vbytes;
      if (gv_s == val_end) gv_s = val_base;
      o += a.out_stride * out_vbytes;
      if (MP == 1) om += a.out_stride;
      if (OBITS && FAST != 2 && (FAST == 1 || ob_row)) ob_row += a.obits_stride;
    }
    cp_async_wait<0>();
  }
}

static size_t win_smem_bytes(int stages, int stage_elems, int stage_melems, bool mask, int nin) {
  return (size_t)stages * nin * ((size_t)stage_elems * 8 + (mask ? stage_melems : 0));
}

template <int K, int MP, int CV, int S, bool CORR, bool AGG, int FAST = 0>
static int launch_window_c(const g2p_plan* p, WinArgs a, cudaStream_t st) {
  constexpr bool MASK = MP == 1 || MP == 2;  // byte masks are the only ones staged
  const size_t smem = win_smem_bytes(S, a.stage_elems, a.stage_melems, MASK, AGG ? 2 : 1);
  auto kern = apply_window_kernel<K, MP, CV, S, CORR, AGG, FAST>;
  G2P_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  G2P_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWinThreads, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t grid_cap = (int64_t)per_sm * num_sms();
  a.stages = S;
  // message chunks: long enough to amortise the per-item index/weight loads, short enough to
  // give every CTA several items
  // (measured on O1280 -> EFAS, tools/chunk_sweep.py: 6 items per CTA against 8 is 1 % faster at 383 messages and
  // 5-8 % faster at 32-128 messages; 4 and fewer lose at 383 to the uneven last round)
  int per_cta = 6;
  if (const char* ev = getenv("G2P_WIN_ITEMS_PER_CTA")) per_cta = std::max(1, atoi(ev));
  int64_t want_chunks = ceil_div(per_cta * grid_cap, p->n_tiles);
  int64_t max_chunks = ceil_div(a.b, 8);
  int64_t n_chunks = std::max<int64_t>(1, std::min(want_chunks, max_chunks));
  a.msgs_per_chunk = (int)ceil_div(a.b, n_chunks);
  n_chunks = ceil_div(a.b, a.msgs_per_chunk);
  a.n_items = (int64_t)p->n_tiles * n_chunks;
  const int grid = (int)std::min<int64_t>(a.n_items, grid_cap);
  kern<<<grid, kWinThreads, smem, st>>>(a);
  G2P_LAUNCHED();
  return G2P_OK;
}

template <int K, int MP, int CV, int S>
static int launch_window_s(const g2p_plan* p, WinArgs a, cudaStream_t st) {
  if (a.agg_items) {
    if (MP >= 3) return set_error(G2P_ERR_UNSUPPORTED, "g2p_apply_planned_fused: byte masks only (PROPAGATE / FILL)");
    // the fused pipeline is instantiated for the window sizes it is used with (CV <= 4: windows up to 2048 values)
    if (CV > 4) return set_error(G2P_ERR_UNSUPPORTED, "g2p_apply_planned_fused: window of %d values is too large; use g2p_manip + g2p_apply_planned", p->max_window);
    if constexpr (CV <= 4 && MP < 3)
      return (a.corr_gem || a.post_on) ? launch_window_c<K, MP, CV, S, true, true>(p, a, st)
                                       : launch_window_c<K, MP, CV, S, false, true>(p, a, st);
  }
  if (a.corr_gem || a.post_on) return launch_window_c<K, MP, CV, S, true, false>(p, a, st);
  if constexpr (MP >= 3) {   // resident fields (marked / bit masks), no epilogue: the switches of the loop at compile time
    const char* ev = getenv("G2P_WIN_NO_FAST");   // experiments / tests: the run-time-switch kernel
    const bool no_fast = ev && atoi(ev);
    if (!a.pairwise && !no_fast) {
      if (!a.out_mask) return launch_window_c<K, MP, CV, S, false, false, 2>(p, a, st);
      if (a.obits_ballot) return launch_window_c<K, MP, CV, S, false, false, 1>(p, a, st);
    }
  }
  return launch_window_c<K, MP, CV, S, false, false>(p, a, st);
}

template <int K, int MP, int CV>
static int launch_window_cv(const g2p_plan* p, WinArgs a, cudaStream_t st) {
  // four stages when they fit in ~72 KB per CTA (three CTAs per SM), else three (the fold pass runs one
  // message ahead of the gathers, so three is the minimum)
  // (measured on O1280 -> EFAS, masks: 3 stages 1.25 ms, 4 stages 1.11 ms, 5 stages 1.16 ms, 6 stages 1.25 ms - more
  // shared memory per CTA leaves less L1 for the table and output traffic)
  const int nin = a.agg_items ? 2 : 1;
  constexpr bool staged_mask = MP == 1 || MP == 2;
  if constexpr (MP == 4) {
    // marked resident fields, no epilogue: the lagged message loop (mbarrier rings, five stages) when its ring fits
    // next to three CTAs per SM.  G2P_WIN_NO_LAG=1 / G2P_WIN_NO_FAST=1: the __syncthreads loop (experiments, tests).
    const char* e1 = getenv("G2P_WIN_NO_FAST");
    const char* e2 = getenv("G2P_WIN_NO_LAG");
    const bool off = (e1 && atoi(e1)) || (e2 && atoi(e2));
    if (!off && !a.agg_items && !a.corr_gem && !a.post_on && !a.pairwise && (!a.out_mask || a.obits_ballot) &&
        win_smem_bytes(5, a.stage_elems, a.stage_melems, false, 1) <= (size_t)72 * 1024)
      return a.out_mask ? launch_window_c<K, MP, CV, 5, false, false, 1>(p, a, st)
                        : launch_window_c<K, MP, CV, 5, false, false, 2>(p, a, st);
  }
  if (win_smem_bytes(4, a.stage_elems, a.stage_melems, staged_mask, nin) <= (size_t)(nin == 2 ? 100 : 72) * 1024)
    return launch_window_s<K, MP, CV, 4>(p, a, st);
  if (win_smem_bytes(3, a.stage_elems, a.stage_melems, staged_mask, nin) > 220 * 1024)
    return set_error(G2P_ERR_UNSUPPORTED, "g2p_apply_planned: window of %d elements does not fit in shared memory", p->max_window);
  return launch_window_s<K, MP, CV, 3>(p, a, st);
}

template <int K, int MP>
static int launch_window(const g2p_plan* p, WinArgs a, cudaStream_t st) {
  a.stage_elems = ((p->max_window + 1 + 15) / 16) * 16;  // +1: the sentinel slot
  a.stage_melems = ((p->max_mwindow + 15) / 16) * 16;
  const int chunks = (p->max_window / 2 + kWinThreads - 1) / kWinThreads;  // 16-byte value chunks per thread
#ifdef G2P_DEV_BUILD   // `make dev`: kernel experiments on the O1280 -> EFAS table only (compiles in a fifth of the time)
  if (chunks != 3) return set_error(G2P_ERR_UNSUPPORTED, "g2p_apply_planned: dev build, CV = 3 only");
  return launch_window_cv<K, MP, 3>(p, a, st);
#endif
  if (chunks <= 2) return launch_window_cv<K, MP, 2>(p, a, st);
  if (chunks <= 3) return launch_window_cv<K, MP, 3>(p, a, st);
  if (chunks <= 4) return launch_window_cv<K, MP, 4>(p, a, st);
  if (chunks <= 8) return launch_window_cv<K, MP, 8>(p, a, st);
  return launch_window_cv<K, MP, 16>(p, a, st);
}

}  // namespace g2p

using namespace g2p;

extern "C" {

int g2p_plan_create(const int32_t* idx, int64_t m, int k, int64_t n, int64_t rows, int64_t cols, g2p_plan** out,
                    void* stream) {
  if (!out) return set_error(G2P_ERR_INVALID, "g2p_plan_create: out is null");
  *out = nullptr;
  if (!idx || m < 1 || k < 1 || k > 16 || n < 1 || n > 0x7fffffffLL)
    return set_error(G2P_ERR_INVALID, "g2p_plan_create: bad arguments");
  if (rows <= 0 || cols <= 0) {
    rows = 1;
    cols = m;
  }
  if (rows * cols != m) return set_error(G2P_ERR_INVALID, "g2p_plan_create: rows*cols != m");
  // the windowed kernel keeps byte offsets of source rows (8*n) in int32 and of output rows (8*m) in uint32
  if (n >= (1LL << 28) || m >= (1LL << 29))
    return set_error(G2P_ERR_UNSUPPORTED, "g2p_plan_create: grids beyond 2^28 source / 2^29 target points use g2p_apply");
  cudaStream_t st = as_stream(stream);
  g2p_plan* p = new g2p_plan();
  p->m = m;
  p->n = n;
  p->n_pad = ceil_div(n, kBlockElems) * kBlockElems;
  p->k = k;
  p->rows = rows;
  p->cols = cols;
  // tile shape: 1024 targets as tile_rows x tile_cols (powers of two).  A flat table is one row; for a 2-D
  // target map the candidates are measured on the table itself and the one with the smallest total
  // window wins (compact tiles suit projected grids whose rows cut across the source rings).
  struct Shape { int tr, tc; };
  std::vector<Shape> cand;
  if (rows == 1) {
    cand.push_back({1, kTileTargets});
  } else {
    for (int tc : {64, 128, 32, 256, 16, 512, 1024}) {
      const int tr = kTileTargets / tc;
      if ((tc / 2 >= cols && tc > 4) || (tr / 2 >= rows && tr > 1)) continue;  // mostly empty tiles
      cand.push_back({tr, tc});
    }
    if (cand.empty()) cand.push_back({1, kTileTargets});
    if (const char* ev = getenv("G2P_WIN_TILE_COLS")) {  // experiments: force one tile shape
      const int tc = atoi(ev);
      if (tc >= 4 && tc <= kTileTargets && (tc & (tc - 1)) == 0) {
        cand.clear();
        cand.push_back({kTileTargets / tc, tc});
      }
    }
  }
  cudaGetDevice(&p->device);
  int32_t* d_status = nullptr;
  unsigned long long* d_total = nullptr;
  cudaError_t e = cudaMalloc(&p->lidx, (size_t)k * m * sizeof(uint16_t));
  if (e == cudaSuccess) e = cudaMalloc(&d_status, 4 * sizeof(int32_t) + 2 * sizeof(unsigned long long));
  if (e != cudaSuccess) {
    g2p_plan_destroy(p);
    return cuda_fail(e, "g2p_plan_create: cudaMalloc", __FILE__, __LINE__);
  }
  d_total = reinterpret_cast<unsigned long long*>(d_status + 4);
  int32_t status[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // + total window (u64) + total gather wavefronts (u64)
  auto run = [&](Shape sh, int want_rows, int dry) -> cudaError_t {
    p->tile_rows = sh.tr;
    p->tile_cols = sh.tc;
    p->tiles_c = (int)ceil_div(cols, sh.tc);
    p->tiles_r = (int)ceil_div(rows, sh.tr);
    p->n_tiles = p->tiles_c * p->tiles_r;
    cudaError_t err = cudaMemsetAsync(d_status, 0, 4 * sizeof(int32_t) + 2 * sizeof(unsigned long long), st);
    if (err != cudaSuccess) return err;
    if (!dry) {
      if (p->ranges) cudaFree(p->ranges);
      if (p->tile_info) cudaFree(p->tile_info);
      p->ranges = nullptr;
      p->tile_info = nullptr;
      err = cudaMalloc(&p->ranges, (size_t)p->n_tiles * kMaxRanges * sizeof(Range));
      if (err == cudaSuccess) err = cudaMalloc(&p->tile_info, (size_t)p->n_tiles * 4 * sizeof(int32_t));
      if (err == cudaSuccess) err = cudaMemsetAsync(p->tile_info, 0, (size_t)p->n_tiles * 4 * sizeof(int32_t), st);
      if (err != cudaSuccess) return err;
    }
    PlanArgs a{};
    a.idx = idx;
    a.m = m;
    a.n = n;
    a.n_pad = p->n_pad;
    a.k = k;
    a.g.rows = rows;
    a.g.cols = cols;
    a.g.tile_rows = sh.tr;
    a.g.tile_cols = sh.tc;
    a.g.tiles_c = p->tiles_c;
    a.g.threads_per_row = sh.tc / kWinTPT;
    {
      // half-warp block: `want_rows` rows when the tile has them, the rest of the 16 lanes along the row
      int br = std::min(want_rows, sh.tr);
      int bc = std::min(16 / br, a.g.threads_per_row);
      br = 16 / bc;
      int lg = 0;
      while ((1 << lg) < bc) ++lg;
      p->block_cols_log2 = lg;
      p->skew_mod = br == 1 ? 2 : 4;
      p->skew_step = br == 1 ? 8 : 4;
      if (const char* e = getenv("G2P_WIN_SKEW_MOD")) {
        const int v = atoi(e);
        if (v == 1 || v == 2 || v == 4) { p->skew_mod = v; p->skew_step = v == 1 ? 0 : 16 / v; }
      }
    }
    a.g.block_cols_log2 = p->block_cols_log2;
    a.skew_mod = p->skew_mod;
    a.skew_step = p->skew_step;
    a.lidx = p->lidx;
    a.ranges = p->ranges;
    a.tile_info = p->tile_info;
    a.status = d_status;
    a.total_window = d_total;
    a.total_wavefronts = d_total + 1;
    a.dry = dry;
    a.swap01 = 2;
    a.swap_sweeps = 2;
    if (const char* ev = getenv("G2P_WIN_SWAP")) a.swap01 = std::max(0, std::min(2, atoi(ev)));   // experiments: 0 = distance order
    if (const char* ev = getenv("G2P_WIN_SWAP_SWEEPS")) a.swap_sweeps = std::max(0, std::min(4, atoi(ev)));
    plan_kernel<<<p->n_tiles, kWinThreads, kPlanSpanBlocks / 8, st>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    err = cudaGetLastError();
    if (err == cudaSuccess) err = cudaMemcpyAsync(status, d_status, sizeof(status), cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess) err = cudaStreamSynchronize(st);
    return err;
  };
  int rc = G2P_OK;
  int best = -1, last_fail = 0;
  unsigned long long best_total = ~0ull;
  if ((int64_t)ceil_div(cols, 4) * rows > 0x7fffffffLL) e = cudaErrorInvalidValue;
  for (size_t i = 0; e == cudaSuccess && cand.size() > 1 && i < cand.size(); ++i) {
    e = run(cand[i], 1, 1);
    if (e != cudaSuccess) break;
    unsigned long long total;
    memcpy(&total, &status[4], sizeof(total));
    if (status[0] == 0 && total < best_total) {
      best_total = total;
      best = (int)i;
    } else if (status[0] != 0) {
      last_fail = status[0];
    }
  }
  if (e == cudaSuccess) {
    if (cand.size() == 1) best = 0;
    if (best < 0) {
      status[0] = last_fail ? last_fail : 1;
    } else {
      // half-warp block shape: the one whose gathers take the fewest shared-memory wavefronts on this table
      // (O1280 -> EFAS: 2x8 blocks 1.41 per half-warp load against 1.62 for 1x16, 0.80 ms against 0.89 ms per 383
      // messages; on global lat/lon targets, whose rows run along the source rings, 1x16 wins).  4x4 blocks
      // have still fewer conflicts on projected grids but their stores touch 4 rows per request - measured equal
      // to 2x8 at best, so they are only used on request (G2P_WIN_BLOCK_ROWS=4).
      int want_rows = 0;
      if (const char* ev = getenv("G2P_WIN_BLOCK_ROWS")) want_rows = atoi(ev);
      if (want_rows != 1 && want_rows != 2 && want_rows != 4) {
        unsigned long long best_waves = ~0ull;
        for (int cand_rows : {1, 2}) {
          if (cand_rows > cand[best].tr) continue;
          e = run(cand[best], cand_rows, 1);
          if (e != cudaSuccess) break;
          unsigned long long waves;
          memcpy(&waves, &status[6], sizeof(waves));
          // a 2-row block splits a warp's store over more rows: it must save 10 % of the wavefronts to pay
          // (GloFAS 0.05 deg: 1.41 against 1.48 per load, yet 0.90 ms against 0.87 ms for 16 messages)
          if (status[0] == 0 && (double)waves < (cand_rows == 1 ? 1.0 : 0.9) * (double)best_waves) {
            best_waves = waves;
            want_rows = cand_rows;
          }
        }
        if (want_rows == 0) want_rows = 1;
      }
      if (e == cudaSuccess) e = run(cand[best], want_rows, 0);
    }
  }
  cudaFree(d_status);
  if (e != cudaSuccess) rc = cuda_fail(e, "g2p_plan_create", __FILE__, __LINE__);
  else if (status[0] != 0) {
    static const char* why[] = {"", "a tile's indexes span more than 2 M source points", "a tile needs more than 96 index ranges",
                                "a tile's window exceeds 8176 source values"};
    rc = set_error(G2P_ERR_UNSUPPORTED, "g2p_plan_create: table is not windowable (%s); use g2p_apply", why[status[0] & 3]);
  }
  if (rc != G2P_OK) {
    g2p_plan_destroy(p);
    return rc;
  }
  {
    unsigned long long total;
    memcpy(&total, &status[4], sizeof(total));
    p->mean_window = (double)total / (double)p->n_tiles;
  }
  p->max_window = status[1];
  p->max_mwindow = status[3];
  p->max_ranges = status[2];
  {
    unsigned long long waves;
    memcpy(&waves, &status[6], sizeof(waves));
    p->gather_wavefronts = (double)waves / ((double)p->n_tiles * (kWinThreads / 16) * kWinTPT * k);
  }
  *out = p;
  return G2P_OK;
}

int g2p_plan_destroy(g2p_plan* p) {
  if (!p) return G2P_OK;
  if (p->lidx) cudaFree(p->lidx);
  if (p->ranges) cudaFree(p->ranges);
  if (p->tile_info) cudaFree(p->tile_info);
  if (p->item_scratch) cudaFree(p->item_scratch);
  delete p;
  return G2P_OK;
}

int g2p_plan_get_info(const g2p_plan* p, g2p_plan_info* info) {
  if (!p || !info) return set_error(G2P_ERR_INVALID, "g2p_plan_get_info: null argument");
  info->m = p->m;
  info->n = p->n;
  info->k = p->k;
  info->tile_rows = p->tile_rows;
  info->tile_cols = p->tile_cols;
  info->n_tiles = p->n_tiles;
  info->max_window = p->max_window;
  info->max_ranges = p->max_ranges;
  info->mean_window = p->mean_window;
  info->block_rows = 16 >> p->block_cols_log2;
  info->block_cols = 1 << p->block_cols_log2;
  info->gather_wavefronts = p->gather_wavefronts;
  info->bytes = (int64_t)p->k * p->m * 2 + (int64_t)p->n_tiles * (kMaxRanges * (int64_t)sizeof(Range) + 16);
  return G2P_OK;
}

static int apply_planned(const g2p_plan* p, const double* coef, const double* src, const uint8_t* src_mask, int b,
                         int64_t src_stride, int64_t out_stride, double mv_out, int mask_policy, const g2p_correction* corr,
                         double* out, uint8_t* out_mask, void* stream, const g2p_manip_desc* fused = nullptr, int b_in = 0,
                         const g2p_writer_post* post = nullptr) {
  if (!p || !src || !out) return set_error(G2P_ERR_INVALID, "g2p_apply_planned: null plan/src/out");
  if (p->k > 1 && !coef) return set_error(G2P_ERR_INVALID, "g2p_apply_planned: coef is required for k > 1");
  if (p->k != 1 && p->k != 4) return set_error(G2P_ERR_UNSUPPORTED, "g2p_apply_planned: k=%d (1 and 4 are planned; use g2p_apply)", p->k);
  if (b < 0) return set_error(G2P_ERR_INVALID, "g2p_apply_planned: b < 0");
  const int pairwise = (mask_policy & G2P_SUM_PAIRWISE) != 0;
  mask_policy &= ~G2P_SUM_PAIRWISE;
  const bool mask = (mask_policy == G2P_MASK_PROPAGATE || mask_policy == G2P_MASK_FILL);   // byte masks
  const bool bits = mask_policy == G2P_MASK_PROPAGATE_BITS, marked = mask_policy == G2P_MASK_MARKED;
  if (!mask && !bits && !marked && mask_policy != G2P_MASK_AS_REFERENCE)
    return set_error(G2P_ERR_INVALID, "g2p_apply_planned: unknown mask_policy %d", mask_policy);
  if ((mask || bits) && !src_mask) return set_error(G2P_ERR_INVALID, "g2p_apply_planned: PROPAGATE / FILL need src_mask");
  if ((mask_policy == G2P_MASK_PROPAGATE || bits) && !out_mask) return set_error(G2P_ERR_INVALID, "g2p_apply_planned: PROPAGATE needs out_mask");
  if (bits && (src_stride % 8) != 0)
    return set_error(G2P_ERR_UNSUPPORTED, "g2p_apply_planned: bit-packed masks need src_stride %% 8 == 0; use g2p_apply");
  if ((bits || marked) && out_mask && (reinterpret_cast<uintptr_t>(out_mask) & 3))
    return set_error(G2P_ERR_INVALID, "g2p_apply_planned: a bit-packed out_mask must be 4-byte aligned");
  if ((bits || marked) && fused) return set_error(G2P_ERR_UNSUPPORTED, "g2p_apply_planned_fused: byte masks only (PROPAGATE / FILL)");
  if (src_stride < p->n_pad || out_stride < p->m) return set_error(G2P_ERR_INVALID, "g2p_apply_planned: strides shorter than rows (src rows must be padded to a multiple of 16)");
  // bulk copies need 16-byte aligned sources: value rows at 16 B, mask rows at 16 B
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (src_stride % 2) != 0)
    return set_error(G2P_ERR_UNSUPPORTED, "g2p_apply_planned: src rows must be 16-byte aligned; use g2p_apply");
  if (mask && ((reinterpret_cast<uintptr_t>(src_mask) & 15) || (src_stride % 16) != 0))
    return set_error(G2P_ERR_UNSUPPORTED, "g2p_apply_planned: mask rows must be 16-byte aligned (stride %% 16 == 0); use g2p_apply");
  if (b == 0) return G2P_OK;
  WinArgs a{};
  a.lidx = p->lidx;
  a.coef = coef;
  a.ranges = p->ranges;
  a.tile_info = p->tile_info;
  a.src = src;
  a.src_mask = src_mask;
  a.out = out;
  a.out_mask = out_mask;
  a.m = p->m;
  a.src_stride = src_stride;
  a.out_stride = out_stride;
  a.sbits_stride = src_stride / 8;
  a.obits_stride = G2P_BITS_ROW_BYTES(out_stride) / 4;
  if ((bits || marked) && out_mask) {
    // whole mask bytes come from warp ballots when every byte of a mask row belongs to one warp and slot: rows start
    // on a byte (one row, or cols % 8 == 0) and a slot's warp footprint is whole 8-column groups
    const int tpr = p->tile_cols / kWinTPT;
    a.obits_ballot = (p->rows == 1 || p->cols % 8 == 0) && tpr % 8 == 0 && p->block_cols_log2 >= 2;
    if (const char* ev = getenv("G2P_WIN_OBITS_ATOMIC")) if (atoi(ev)) a.obits_ballot = 0;
    if (!a.obits_ballot)   // only masked outputs touch the bit rows (atomicOr), everything else is this memset
      G2P_CUDA(cudaMemsetAsync(out_mask, 0, (size_t)b * a.obits_stride * 4, as_stream(stream)));
  }
  a.g.rows = p->rows;
  a.g.cols = p->cols;
  a.g.tile_rows = p->tile_rows;
  a.g.tile_cols = p->tile_cols;
  a.g.tiles_c = p->tiles_c;
  a.g.threads_per_row = p->tile_cols / kWinTPT;
  a.g.block_cols_log2 = p->block_cols_log2;
  a.b = b;
  a.pairwise = pairwise;
  a.n_tiles = p->n_tiles;
  a.mv_out = mv_out;
  AggItem* d_items = nullptr;
  if (fused) {
    // b = number of OUTPUT messages (fused->n_out), src holds b_in input messages
    if (fused->n_out != b || b_in < 1 || !fused->items) return set_error(G2P_ERR_INVALID, "g2p_apply_planned_fused: bad item list");
    std::vector<AggItem> items(b);
    for (int o = 0; o < b; ++o) {
      const g2p_manip_item& it = fused->items[o];
      const bool one = it.kind == G2P_AGG_PICK || it.kind == G2P_AGG_COPY;
      if (it.kind < G2P_AGG_PICK || it.kind > G2P_AGG_COPY || it.n_in < 1 || it.n_in > 2 || (one && it.n_in != 1) ||
          (it.kind == G2P_AGG_ACCUM && it.n_in != 2))
        return set_error(G2P_ERR_UNSUPPORTED, "g2p_apply_planned_fused: item %d (kind %d, %d inputs): at most two inputs per output; use g2p_manip", o, it.kind, it.n_in);
      if (mask && it.mask_all) return set_error(G2P_ERR_UNSUPPORTED, "g2p_apply_planned_fused: mask_all items need g2p_manip");
      AggItem& d = items[o];
      d.kind = it.kind;
      d.in0 = it.in[0];
      d.in1 = it.n_in > 1 ? it.in[1] : -1;
      d.pad = 0;
      d.unit_time = it.unit_time;
      d.divisor = it.divisor;
      if (d.in0 >= b_in || d.in1 >= b_in || (d.in0 < 0 && it.kind != G2P_AGG_PICK) || (d.in1 < 0 && it.kind == G2P_AGG_AVERAGE && it.n_in > 1))
        return set_error(G2P_ERR_INVALID, "g2p_apply_planned_fused: item %d references a message outside 0..%d", o, b_in - 1);
    }
    const size_t need = items.size() * sizeof(AggItem);
    if (p->item_scratch_bytes < need) {   // one plan is used from one stream at a time (as its table is)
      if (p->item_scratch) cudaFree(p->item_scratch);
      p->item_scratch = nullptr;
      p->item_scratch_bytes = 0;
      G2P_CUDA(cudaMalloc(&p->item_scratch, need * 2));
      p->item_scratch_bytes = need * 2;
    }
    d_items = reinterpret_cast<AggItem*>(p->item_scratch);
    // pageable source: the copy is staged before the call returns, `items` may go out of scope
    G2P_CUDA(cudaMemcpyAsync(d_items, items.data(), need, cudaMemcpyHostToDevice, as_stream(stream)));
    a.agg_items = d_items;
    a.op1 = fused->conv_op1;
    a.op2 = fused->conv_op2;
    a.c1 = fused->conv_c1;
    a.c2 = fused->conv_c2;
    a.mv_grib = fused->mv_grib;
    a.cut_off = fused->cut_off_negative;
  }
  if (corr) {
    if (!corr->gem || !corr->dem_term || !corr->valid) return set_error(G2P_ERR_INVALID, "g2p_apply_planned_corrected: null correction arrays");
    if (corr->mode != 0 && corr->mode != 1) return set_error(G2P_ERR_INVALID, "g2p_apply_planned_corrected: unknown correction mode %d", corr->mode);
    a.corr_gem = corr->gem;
    a.corr_dem_term = corr->dem_term;
    a.corr_valid = corr->valid;
    a.corr_mv = corr->mv;
    a.corr_mode = corr->mode;
  }
  if (post) {
    a.post_on = 1;
    a.post_clone = post->clone_mask;
    a.post_use_match = post->use_match;
    a.post_mv_match = post->mv_match;
    a.post_mv_write = post->mv_write;
    a.post_f32 = post->out_f32 ? 1 : 0;
  }
  cudaStream_t st = as_stream(stream);
  int rc;
  if (p->k == 1) {
    rc = bits ? launch_window<1, 3>(p, a, st) : marked ? launch_window<1, 4>(p, a, st)
         : mask_policy == G2P_MASK_PROPAGATE ? launch_window<1, 1>(p, a, st) : mask ? launch_window<1, 2>(p, a, st) : launch_window<1, 0>(p, a, st);
  } else {
    rc = bits ? launch_window<4, 3>(p, a, st) : marked ? launch_window<4, 4>(p, a, st)
         : mask_policy == G2P_MASK_PROPAGATE ? launch_window<4, 1>(p, a, st) : mask ? launch_window<4, 2>(p, a, st) : launch_window<4, 0>(p, a, st);
  }
  return rc;
}

int g2p_apply_planned(const g2p_plan* p, const double* coef, const double* src, const uint8_t* src_mask, int b,
                      int64_t src_stride, int64_t out_stride, double mv_out, int mask_policy, double* out,
                      uint8_t* out_mask, void* stream) {
  return apply_planned(p, coef, src, src_mask, b, src_stride, out_stride, mv_out, mask_policy, nullptr, out, out_mask, stream);
}

int g2p_apply_planned_corrected(const g2p_plan* p, const double* coef, const double* src, const uint8_t* src_mask, int b,
                                int64_t src_stride, int64_t out_stride, double mv_out, int mask_policy,
                                const g2p_correction* corr, double* out, uint8_t* out_mask, void* stream) {
  if (!corr) return set_error(G2P_ERR_INVALID, "g2p_apply_planned_corrected: corr is null");
  return apply_planned(p, coef, src, src_mask, b, src_stride, out_stride, mv_out, mask_policy, corr, out, out_mask, stream);
}

int g2p_apply_planned_fused(const g2p_plan* p, const double* coef, const g2p_manip_desc* desc, const double* src,
                            const uint8_t* src_mask, int b_in, int64_t src_stride, int64_t out_stride, double mv_out,
                            int mask_policy, const g2p_correction* corr, double* out, uint8_t* out_mask, void* stream) {
  if (!desc) return set_error(G2P_ERR_INVALID, "g2p_apply_planned_fused: desc is null");
  for (int op : {desc->conv_op1, desc->conv_op2})
    if (op < G2P_OP_NONE || op > G2P_OP_DIV) return set_error(G2P_ERR_INVALID, "g2p_apply_planned_fused: unknown conversion op %d", op);
  return apply_planned(p, coef, src, src_mask, desc->n_out, src_stride, out_stride, mv_out, mask_policy, corr, out, out_mask,
                       stream, desc, b_in);
}

int g2p_apply_planned_post(const g2p_plan* p, const double* coef, const double* src, const uint8_t* src_mask, int b,
                           int64_t src_stride, int64_t out_stride, double mv_out, int mask_policy, const g2p_correction* corr,
                           const g2p_writer_post* post, double* out, uint8_t* out_mask, void* stream) {
  return apply_planned(p, coef, src, src_mask, b, src_stride, out_stride, mv_out, mask_policy, corr, out, out_mask, stream,
                       nullptr, 0, post);
}

}  // extern "C"
