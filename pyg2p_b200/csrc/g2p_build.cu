// K1-K5: lat/lon -> xyz, cube-sphere source index, k-NN search, min_upper_bound reduce and
// nearest / invdist weights for sm_100a.
//
// Replaces ScipyInterpolation.__init__ / interpolate_split / to_3d / _build_nn /
// _build_weights_invdist (reference src/pyg2p/main/interpolation/scipy_interpolation_lib.py
// :523-570, :615-663, :665-696, :698-727, :774-812 + :1018-1026) and the scipy cKDTree they
// call.  Design:
//   * source points are binned into an equi-angular cube-sphere grid (6 faces x G x G cells),
//     counting-sorted by cell key into 32-byte records {x, y, z, id} so one candidate is one
//     32-byte sector and one row of cells is one contiguous range of records;
//   * a query is a cap search: for a radius D the cap {p : |p-q| <= D} is bounded, on each
//     face it can reach, by an (alpha, beta) rectangle computed from the tangent planes through
//     the face axes; all cells of the rectangles are scanned by a group of LANES threads (2 by
//     default) with a register-resident sorted top-k per lane plus the best distance behind it
//     (for the tie flags), merged with group shuffles;
//   * the search ends when the k-th best distance is <= D, so every point that could tie or
//     beat it has been examined; otherwise D is set to the k-th best (or doubled);
//   * squared distances are ((dx*dx + dy*dy) + dz*dz) in float64 with separate IEEE ops - the
//     cKDTree arithmetic - and ordered by (d2, source id);
//   * lat/lon -> xyz uses double-double sin/cos so that the coordinates equal libm's bit for bit in
//     99.7 % of the points.
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <mutex>
#include <vector>

#include "g2p_common.cuh"
#include "g2p_sincosf.h"

namespace g2p {

constexpr double kPi = 3.14159265358979323846;
constexpr double kQuarterPi = kPi / 4.0;
constexpr int kMaxG = 2896;  // 6*G*G <= ~50.3M cells

struct __align__(32) Pt {
  double x, y, z;
  int32_t id;
  int32_t pad;
};

// one candidate record = one 32-byte sector = one 256-bit read-only load (LDG.E.256 on sm_100a)
// a block of kFarB x kFarB cells of one face that holds source points, with the cap (unit centre, angular radius) that
// bounds it: the coarse level of the search for queries far from every source point
constexpr int kFarB = 16;
struct __align__(32) FarBlock {
  float cx, cy, cz;     // unit vector to the middle of the block
  float cos_r, sin_r;   // of the angular radius (a little enlarged)
  int32_t f, ib0, ia0;  // face, first cell row / column
};

__device__ __forceinline__ void load_pt(const Pt* p, double& x, double& y, double& z, int& id) {
  long long w;
  asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=l"(w) : "l"(p));
  id = (int)(w & 0xffffffffLL);
}

struct Rot {
  int on;
  int libm_fma;  // float32 maps: which glibc build of sinf / cosf is restated (g2p_sincosf.h); 1 = the FMA build
  double xx, xy, xz, yx, yy, yz, zx, zz;  // x = (xx*ux + xy*uy) + xz*uz ; y = (yx*ux + yy*uy) - yz*uz ; z = zx*ux + zz*uz
};

static Rot make_rot(const g2p_rotation* r) {
  Rot o{};
  // G2P_LIBM_FMA=0: hosts whose glibc resolves sinf / cosf to the non-FMA (sse2) build
  o.libm_fma = 1;
  if (const char* ev = getenv("G2P_LIBM_FMA")) o.libm_fma = atoi(ev) != 0;
  if (!r) return o;
  // teta = -radians(90 + latSP), fi = -radians(lonSP)  (scipy_interpolation_lib.py:674-675)
  const double deg = kPi / 180.0;
  double teta = -((90 + r->lat_south_pole_deg) * deg);
  double fi = -(r->lon_south_pole_deg * deg);
  o.on = 1;
  o.xx = cos(teta) * cos(fi);
  o.xy = sin(fi);
  o.xz = sin(teta) * cos(fi);
  o.yx = cos(teta) * sin(fi);
  o.yy = cos(fi);
  o.yz = sin(teta) * sin(fi);
  o.zx = -sin(teta);
  o.zz = cos(teta);
  return o;
}

// ------------------------------------------------------------------ correctly rounded sin / cos
// libdevice's float64 sin/cos are accurate to 1-2 ulp; the reference evaluates `to_3d` with libm (numexpr /
// numpy), whose results are correctly rounded in ~99.9 % of the cases.  With libdevice only 67 % of the
// source / target points get bit-identical xyz; evaluating sin and cos in double-double arithmetic
// (Cody-Waite reduction by pi/2 in three 33-bit parts, Taylor series with the leading terms in
// double-double, ~2^-75 relative error before the final rounding) reproduces libm in 99.86 % of the values,
// so that distances, weights and near-tie decisions follow the reference's bit for bit almost everywhere.
static __constant__ double kSinC[11][2] = {
    {-0x1.5555555555555p-3, -0x1.5555555555555p-57},
    {0x1.1111111111111p-7, 0x1.1111111111111p-63},
    {-0x1.a01a01a01a01ap-13, -0x1.a01a01a01a01ap-73},
    {0x1.71de3a556c734p-19, -0x1.c154f8ddc6c00p-73},
    {-0x1.ae64567f544e4p-26, 0x1.c062e06d1f209p-80},
    {0x1.6124613a86d09p-33, 0x1.f28e0cc748ebep-87},
    {-0x1.ae7f3e733b81fp-41, -0x1.1d8656b0ee8cbp-97},
    {0x1.952c77030ad4ap-49, 0x1.ac981465ddc6cp-103},
    {-0x1.2f49b46814157p-57, -0x1.2650f61dbdcb4p-112},
    {0x1.71b8ef6dcf572p-66, -0x1.d043ae40c4647p-120},
    {-0x1.761b41316381ap-75, 0x1.3423c7d91404fp-130}};
static __constant__ double kCosC[11][2] = {
    {-0x1.0000000000000p-1, 0x0.0p+0},
    {0x1.5555555555555p-5, 0x1.5555555555555p-59},
    {-0x1.6c16c16c16c17p-10, 0x1.f49f49f49f49fp-65},
    {0x1.a01a01a01a01ap-16, 0x1.a01a01a01a01ap-76},
    {-0x1.27e4fb7789f5cp-22, -0x1.cbbc05b4fa99ap-76},
    {0x1.1eed8eff8d898p-29, -0x1.2aec959e14c06p-83},
    {-0x1.93974a8c07c9dp-37, -0x1.05d6f8a2efd1fp-92},
    {0x1.ae7f3e733b81fp-45, 0x1.1d8656b0ee8cbp-101},
    {-0x1.6827863b97d97p-53, -0x1.eec01221a8b0bp-107},
    {0x1.e542ba4020225p-62, 0x1.ea72b4afe3c2fp-120},
    {-0x1.0ce396db7f853p-70, 0x1.aebcdbd20331cp-124}};

struct DD {
  double h, l;
};
__device__ __forceinline__ DD fast_two_sum(double a, double b) {
  const double s = __dadd_rn(a, b);
  return {s, __dsub_rn(b, __dsub_rn(s, a))};
}
__device__ __forceinline__ DD two_sum(double a, double b) {
  const double s = __dadd_rn(a, b);
  const double bb = __dsub_rn(s, a);
  return {s, __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb))};
}
__device__ __forceinline__ DD dd_mul(DD a, DD b) {
  const double p = __dmul_rn(a.h, b.h);
  double e = __fma_rn(a.h, b.h, -p);
  e = __dadd_rn(e, __dadd_rn(__dmul_rn(a.h, b.l), __dmul_rn(a.l, b.h)));
  return fast_two_sum(p, e);
}
__device__ __forceinline__ DD dd_add(DD a, DD b) {
  DD s = two_sum(a.h, b.h);
  return fast_two_sum(s.h, __dadd_rn(s.l, __dadd_rn(a.l, b.l)));
}

__device__ __noinline__ void sincos_cr(double x, double* sn, double* cs) {
  if (!(fabs(x) < 1.0e5)) {  // far outside lat/lon ranges (garbage coordinates of MV cells), NaN, inf
    sincos(x, sn, cs);
    return;
  }
  if (x == 0.0) {
    *sn = x;
    *cs = 1.0;
    return;
  }
  const double k = rint(__dmul_rn(x, 0.6366197723675814));
  // r = x - k*pi/2 with pi/2 = C1 + C2 + C3 + C3t, each Ci of 33 bits: k*Ci is exact for |k| < 2^20
  const double t1 = __dsub_rn(x, __dmul_rn(k, 0x1.921fb54400000p+0));
  DD s1 = two_sum(t1, -__dmul_rn(k, 0x1.0b4611a600000p-34));
  DD s2 = two_sum(s1.h, -__dmul_rn(k, 0x1.3198a2e000000p-69));
  const double lo = __dsub_rn(__dadd_rn(s1.l, s2.l), __dmul_rn(k, 0x1.b839a252049c1p-104));
  const DD r = fast_two_sum(s2.h, lo);
  const DD u = dd_mul(r, r);
  // sin r = r + r*u*(S1 + u*(S2 + ...)), cos r = 1 + u*(C1 + u*(C2 + ...)); high orders in double
  double q = kSinC[10][0];
#pragma unroll
  for (int i = 9; i >= 4; --i) q = __fma_rn(u.h, q, kSinC[i][0]);
  DD t = dd_add(dd_mul(u, DD{q, 0.0}), DD{kSinC[3][0], kSinC[3][1]});
#pragma unroll
  for (int i = 2; i >= 0; --i) t = dd_add(dd_mul(u, t), DD{kSinC[i][0], kSinC[i][1]});
  const DD s = dd_add(r, dd_mul(r, dd_mul(u, t)));
  q = kCosC[10][0];
#pragma unroll
  for (int i = 9; i >= 5; --i) q = __fma_rn(u.h, q, kCosC[i][0]);
  t = dd_add(dd_mul(u, DD{q, 0.0}), DD{kCosC[4][0], kCosC[4][1]});
#pragma unroll
  for (int i = 3; i >= 0; --i) t = dd_add(dd_mul(u, t), DD{kCosC[i][0], kCosC[i][1]});
  const DD c = dd_add(DD{1.0, 0.0}, dd_mul(u, t));
  const int n = ((int)k) & 3;
  *sn = (n == 0) ? s.h : (n == 1) ? c.h : (n == 2) ? -s.h : -c.h;
  *cs = (n == 0) ? c.h : (n == 1) ? -s.h : (n == 2) ? -c.h : s.h;
}

// float32 sin / cos exactly as glibc's sinf / cosf return them (what numexpr's float32 opcodes call), widened
__device__ __forceinline__ void sincosf_libm(float ang, int fma_build, double& s, double& c) {
  if (fma_build) {
    s = (double)g2p_scf::sincosf_glibc<true, false>(ang);
    c = (double)g2p_scf::sincosf_glibc<true, true>(ang);
  } else {
    s = (double)g2p_scf::sincosf_glibc<false, false>(ang);
    c = (double)g2p_scf::sincosf_glibc<false, true>(ang);
  }
}

// ------------------------------------------------------------------ K1  lat/lon -> xyz
// f64: radians = deg * (pi/180) in double (numpy deg2rad), libdevice sin/cos, products
// (r*cos(lon))*cos(lat) left to right with separate multiplies.
// f32: radians in float32, sin/cos = glibc's sinf/cosf bit for bit, promoted to double for the products with the double
// radius, result rounded to float32 (reference :691-694) and widened again for the search.
// PAIRED: the two lanes of a lane pair (same query) share the trigonometry - the even lane evaluates the
// longitude, the odd lane the latitude, and they swap results with one shuffle - which halves the cost of the
// double-double sin/cos in the k-NN kernel (pass the group's lane mask; 0 = every thread on its own)
template <bool F32>
__device__ __forceinline__ void latlon_to_xyz(const void* lon, const void* lat, int64_t i, double radius,
                                              const Rot& rot, double& x, double& y, double& z, unsigned pair_mask = 0u) {
  double cx, sx, cy, sy;
  if (pair_mask != 0u) {
    const bool odd = threadIdx.x & 1;
    double s1, c1;
    if (F32) {
      const float dr = 0.017453292519943295f;
      const float ang = __fmul_rn(reinterpret_cast<const float*>(odd ? lat : lon)[i], dr);
      sincosf_libm(ang, rot.libm_fma, s1, c1);
    } else {
      const double ang = __dmul_rn(reinterpret_cast<const double*>(odd ? lat : lon)[i], 0.017453292519943295);
      sincos_cr(ang, &s1, &c1);
    }
    const double s2 = __shfl_xor_sync(pair_mask, s1, 1), c2 = __shfl_xor_sync(pair_mask, c1, 1);
    sx = odd ? s2 : s1;
    cx = odd ? c2 : c1;
    sy = odd ? s1 : s2;
    cy = odd ? c1 : c2;
  } else if (F32) {
    const float dr = 0.017453292519943295f;  // float32(pi/180): numpy's deg2radf constant
    float lo = __fmul_rn(reinterpret_cast<const float*>(lon)[i], dr);
    float la = __fmul_rn(reinterpret_cast<const float*>(lat)[i], dr);
    // numexpr evaluates float32 sin/cos with libm's sinf/cosf: restated bit for bit (g2p_sincosf.h)
    sincosf_libm(lo, rot.libm_fma, sx, cx);
    sincosf_libm(la, rot.libm_fma, sy, cy);
  } else {
    const double dr = 0.017453292519943295;
    double lo = __dmul_rn(reinterpret_cast<const double*>(lon)[i], dr);
    double la = __dmul_rn(reinterpret_cast<const double*>(lat)[i], dr);
    sincos_cr(lo, &sx, &cx);
    sincos_cr(la, &sy, &cy);
  }
  if (!rot.on) {
    x = __dmul_rn(__dmul_rn(radius, cx), cy);
    y = __dmul_rn(__dmul_rn(radius, sx), cy);
    z = __dmul_rn(radius, sy);
  } else {
    double ux, uy, uz = sy;
    if (F32) {  // cos(lons)*cos(lats) is a float32 product in numexpr before meeting the double constants
      ux = (double)__fmul_rn((float)cx, (float)cy);
      uy = (double)__fmul_rn((float)sx, (float)cy);
    } else {
      ux = __dmul_rn(cx, cy);
      uy = __dmul_rn(sx, cy);
    }
    x = __dadd_rn(__dadd_rn(__dmul_rn(rot.xx, ux), __dmul_rn(rot.xy, uy)), __dmul_rn(rot.xz, uz));
    y = __dsub_rn(__dadd_rn(__dmul_rn(rot.yx, ux), __dmul_rn(rot.yy, uy)), __dmul_rn(rot.yz, uz));
    z = __dadd_rn(__dmul_rn(rot.zx, ux), __dmul_rn(rot.zz, uz));
  }
  if (F32) {
    x = (double)(float)x;
    y = (double)(float)y;
    z = (double)(float)z;
  }
}

template <bool F32>
__global__ void xyz_kernel(const void* __restrict__ lon, const void* __restrict__ lat, int64_t n, double radius,
                           Rot rot, double* __restrict__ xyz) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double x, y, z;
    latlon_to_xyz<F32>(lon, lat, i, radius, rot, x, y, z);
    xyz[3 * i + 0] = x;
    xyz[3 * i + 1] = y;
    xyz[3 * i + 2] = z;
  }
}

// ------------------------------------------------------------------ cube-sphere cells
// face f: axis = f>>1, sign = +1 (f even) / -1 (f odd); local c = sign*p[axis] (> 0 on the face),
// a = p[(axis+1)%3], b = p[(axis+2)%3]; alpha = atan2(a, c), beta = atan2(b, c) in [-pi/4, pi/4].
__device__ __forceinline__ void face_local(int f, double x, double y, double z, double& a, double& b, double& c) {
  const int ax = f >> 1;
  const double sg = (f & 1) ? -1.0 : 1.0;
  if (ax == 0) {
    c = sg * x; a = y; b = z;
  } else if (ax == 1) {
    c = sg * y; a = z; b = x;
  } else {
    c = sg * z; a = x; b = y;
  }
}

__device__ __forceinline__ int dominant_face(double x, double y, double z) {
  const double fx = fabs(x), fy = fabs(y), fz = fabs(z);
  if (fx >= fy && fx >= fz) return x >= 0 ? 0 : 1;
  if (fy >= fz) return y >= 0 ? 2 : 3;
  return z >= 0 ? 4 : 5;
}

__device__ __forceinline__ int cell_coord(double ang, int G) {
  int i = (int)floor((ang + kQuarterPi) * ((double)G / (kPi / 2.0)));
  return i < 0 ? 0 : (i >= G ? G - 1 : i);
}

__device__ __forceinline__ uint32_t cell_key(double x, double y, double z, int G) {
  const int f = dominant_face(x, y, z);
  double a, b, c;
  face_local(f, x, y, z, a, b, c);
  const int ia = cell_coord(atan2(a, c), G);
  const int ib = cell_coord(atan2(b, c), G);
  return (uint32_t)((f * G + ib) * G + ia);
}

__device__ __forceinline__ void atomic_min_pos_double(double* addr, double v) {
  atomicMin(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}
__device__ __forceinline__ void atomic_max_pos_double(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// stats[0] = min |p|, stats[1] = max |p| ; coarse occupancy flags for the density estimate
__global__ void stats_kernel(const double* __restrict__ xyz, int64_t n, int Gc, double* __restrict__ stats,
                             uint8_t* __restrict__ coarse) {
  double mn = DBL_MAX, mx = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    const double r = sqrt(x * x + y * y + z * z);
    if (r == r) {
      mn = fmin(mn, r);
      mx = fmax(mx, r);
      coarse[cell_key(x, y, z, Gc)] = 1;
    } else {
      mx = __longlong_as_double(0x7ff0000000000000LL);  // NaN coordinates poison the stats -> rejected by the host
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomic_min_pos_double(&stats[0], mn);
    atomic_max_pos_double(&stats[1], mx);
  }
}

__global__ void key_hist_kernel(const double* __restrict__ xyz, int64_t n, int G, uint32_t* __restrict__ keys,
                                uint32_t* __restrict__ counts) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t key = cell_key(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], G);
    keys[i] = key;
    atomicAdd(&counts[key], 1u);
  }
}

// exclusive scan of uint32 counts, three kernels: per-block scan, scan of block sums, add-back
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanBlock = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads) scan_blocks_kernel(const uint32_t* __restrict__ in,
                                                                   uint32_t* __restrict__ out, int64_t n,
                                                                   uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t warp_sums[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanBlock + (int64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t sum = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0u;
    sum += v[i];
  }
  uint32_t incl = sum;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = (lane < kScanThreads / 32) ? warp_sums[lane] : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    if (lane < kScanThreads / 32) warp_sums[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  uint32_t excl = incl - sum + (warp > 0 ? warp_sums[warp - 1] : 0u);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = excl;
    excl += v[i];
  }
  if (threadIdx.x == kScanThreads - 1) block_sums[blockIdx.x] = warp_sums[kScanThreads / 32 - 1];
}

__global__ void __launch_bounds__(1024) scan_sums_kernel(uint32_t* __restrict__ block_sums, int nb) {
  // one CTA; each thread owns a contiguous slice
  __shared__ uint32_t part[1024];
  const int per = (nb + 1023) / 1024;
  const int b0 = threadIdx.x * per;
  uint32_t s = 0;
  for (int i = b0; i < b0 + per && i < nb; ++i) s += block_sums[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int i = 0; i < 1024; ++i) {
      uint32_t t = part[i];
      part[i] = run;
      run += t;
    }
  }
  __syncthreads();
  uint32_t run = part[threadIdx.x];
  for (int i = b0; i < b0 + per && i < nb; ++i) {
    uint32_t t = block_sums[i];
    block_sums[i] = run;
    run += t;
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_add_kernel(uint32_t* __restrict__ out, int64_t n,
                                                                const uint32_t* __restrict__ block_sums,
                                                                uint32_t* __restrict__ cursor, uint32_t total) {
  const int64_t base = (int64_t)blockIdx.x * kScanBlock + (int64_t)threadIdx.x * kScanItems;
  const uint32_t add = block_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) {
      const uint32_t v = out[base + i] + add;
      out[base + i] = v;
      if (cursor) cursor[base + i] = v;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = total;
}

__global__ void scatter_kernel(const double* __restrict__ xyz, const uint32_t* __restrict__ keys, int64_t n,
                               uint32_t* __restrict__ cursor, Pt* __restrict__ pts) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t pos = atomicAdd(&cursor[keys[i]], 1u);
    Pt p;
    p.x = xyz[3 * i];
    p.y = xyz[3 * i + 1];
    p.z = xyz[3 * i + 2];
    p.id = (int32_t)i;
    p.pad = 0;
    pts[pos] = p;
  }
}

}  // namespace g2p

namespace g2p {
// one thread per block of kFarB x kFarB cells: occupied blocks are appended (in any order) with their bounding cap
__global__ void far_blocks_kernel(const uint32_t* __restrict__ cell_start, int G, int Gb, FarBlock* __restrict__ out,
                                  int* __restrict__ count) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 6 * Gb * Gb) return;
  const int f = t / (Gb * Gb), bb = (t / Gb) % Gb, ba = t % Gb;
  const int ia0 = ba * kFarB, ib0 = bb * kFarB;
  const int ia1 = min(ia0 + kFarB, G), ib1 = min(ib0 + kFarB, G);
  bool occupied = false;
  for (int r = ib0; r < ib1 && !occupied; ++r) {
    const int64_t row = ((int64_t)f * G + r) * G;
    occupied = cell_start[row + ia0] != cell_start[row + ia1];
  }
  if (!occupied) return;
  const double step = (kPi / 2.0) / (double)G;
  auto dir = [&](double al, double be, double& x, double& y, double& z) {
    // inverse of face_local for the point (tan alpha, tan beta, 1) of the face
    const double a = tan(al), b = tan(be), inv = 1.0 / sqrt(a * a + b * b + 1.0);
    const double la = a * inv, lb = b * inv, lc = inv;
    const int ax = f >> 1;
    const double sg = (f & 1) ? -1.0 : 1.0;
    if (ax == 0) { x = sg * lc; y = la; z = lb; }
    else if (ax == 1) { y = sg * lc; z = la; x = lb; }
    else { z = sg * lc; x = la; y = lb; }
  };
  const double a0 = -kQuarterPi + ia0 * step, a1 = -kQuarterPi + ia1 * step;
  const double b0 = -kQuarterPi + ib0 * step, b1 = -kQuarterPi + ib1 * step;
  double cx, cy, cz;
  dir(0.5 * (a0 + a1), 0.5 * (b0 + b1), cx, cy, cz);
  double min_dot = 1.0;
  for (int q = 0; q < 4; ++q) {   // the farthest point of a gnomonic rectangle from an inner point is a corner
    double x, y, z;
    dir((q & 1) ? a1 : a0, (q & 2) ? b1 : b0, x, y, z);
    min_dot = fmin(min_dot, x * cx + y * cy + z * cz);
  }
  const double rad = acos(fmax(-1.0, fmin(1.0, min_dot))) + 2e-4;   // slack >> float rounding of the stored cap
  FarBlock fb;
  fb.cx = (float)cx; fb.cy = (float)cy; fb.cz = (float)cz;
  fb.cos_r = (float)cos(rad); fb.sin_r = (float)sin(rad);
  fb.f = f; fb.ib0 = ib0; fb.ia0 = ia0;
  out[atomicAdd(count, 1)] = fb;
}

}  // namespace g2p

struct g2p_index {
  int device;
  int64_t n;
  int G;
  int64_t n_cells;
  double radius;        // mean of min/max norm
  double user_radius;   // radius given to g2p_index_create (to_3d's `r`), 0 if built from xyz
  double mean_spacing;  // h
  double* xyz;          // (n,3) original order
  g2p::Pt* pts;         // counting-sorted by cell key
  uint32_t* cell_start; // n_cells + 1
  g2p::FarBlock* blocks; // occupied blocks of 16 x 16 cells with their bounding caps (the far search)
  int n_blocks;
  size_t blocks_bytes;
  int64_t bytes;
};

namespace g2p {

static int grid_for(int64_t n, int threads, int waves = 8) {
  int64_t blocks = ceil_div(n, threads);
  int64_t cap = (int64_t)num_sms() * waves;
  return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

// Device buffers of the index and its scratch are recycled through a small per-device free list: a table build
// allocates ~450 MB in nine pieces, and both cudaMalloc and cudaMallocAsync (whose pool regrows unpredictably:
// 0.02 - 55 ms measured for the same sequence) cost more than the 0.75 ms of kernels that build the index.
// Buffers come back when an index is destroyed (after a device synchronisation) or when build_cells is done
// with its scratch (same stream, so stream order protects the reuse).  Index calls on one device are expected
// on one stream at a time.
struct CachedBuf {
  void* p;
  size_t bytes;
};
static std::vector<CachedBuf> g_free_bufs[64];
static std::mutex g_free_mutex;
constexpr size_t kCacheMaxEntries = 40;
constexpr size_t kCacheMaxBytes = (size_t)2 << 30;   // at most 2 GiB parked per device

template <typename T>
static cudaError_t pool_alloc(T** p, size_t bytes, cudaStream_t st) {
  if (bytes == 0) bytes = 1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64) {
    std::lock_guard<std::mutex> lock(g_free_mutex);
    auto& fl = g_free_bufs[dev];
    int best = -1;
    for (int i = 0; i < (int)fl.size(); ++i)
      if (fl[i].bytes >= bytes && fl[i].bytes <= bytes + bytes / 2 + 4096 && (best < 0 || fl[i].bytes < fl[best].bytes)) best = i;
    if (best >= 0) {
      *p = reinterpret_cast<T*>(fl[best].p);
      fl.erase(fl.begin() + best);
      return cudaSuccess;
    }
  }
  (void)st;
  return cudaMalloc(reinterpret_cast<void**>(p), bytes);
}

static void pool_free(void* p, size_t bytes) {
  if (!p) return;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64) {
    std::lock_guard<std::mutex> lock(g_free_mutex);
    auto& fl = g_free_bufs[dev];
    size_t held = 0;
    for (const CachedBuf& b : fl) held += b.bytes;
    if (bytes <= kCacheMaxBytes) {
      // the newest buffer stays, the oldest ones go: after a change of problem size (another grid) the list turns over
      // once and the new sizes are served from it again (without this, a full list of stale sizes made every later
      // build pay cudaMalloc + cudaFree per buffer: 1.16 instead of 0.47 ms for an O320 -> EFAS build)
      std::vector<void*> evict;
      while (!fl.empty() && (fl.size() >= kCacheMaxEntries || held + bytes > kCacheMaxBytes)) {
        held -= fl.front().bytes;
        evict.push_back(fl.front().p);
        fl.erase(fl.begin());
      }
      fl.push_back({p, bytes ? bytes : 1});
      for (void* e : evict) cudaFree(e);
      return;
    }
  }
  cudaFree(p);
}

static void free_index(g2p_index* ix) {
  if (!ix) return;
  // wait for every stream (a query on any of them may still be reading the index), then hand the memory
  // back to the pool - cudaFree would return it to the driver and the next build would pay for it again
  cudaDeviceSynchronize();
  pool_free(ix->xyz, (size_t)ix->n * 24);
  pool_free(ix->pts, (size_t)ix->n * sizeof(g2p::Pt));
  pool_free(ix->cell_start, (size_t)(ix->n_cells + 1) * 4);
  if (ix->blocks) pool_free(ix->blocks, ix->blocks_bytes);
  delete ix;
}

static double env_double(const char* name, double dflt) {
  const char* s = getenv(name);
  if (!s || !*s) return dflt;
  char* end = nullptr;
  double v = strtod(s, &end);
  return (end && end != s) ? v : dflt;
}

// builds cells from ix->xyz (already on device)
static double now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static int build_cells(g2p_index* ix, cudaStream_t st) {
  const int64_t n = ix->n;
  const bool timing = getenv("G2P_TIMING") != nullptr;
  double t0 = now_ms();
  if (timing) cudaStreamSynchronize(st);
  const double t_enter = now_ms();
  // 1. norm range + coarse occupancy (density estimate for regional sources)
  int Gc = (int)floor(sqrt((double)n / 96.0));
  Gc = Gc < 1 ? 1 : (Gc > 64 ? 64 : Gc);
  const int n_coarse = 6 * Gc * Gc;
  double* d_stats = nullptr;
  uint8_t* d_coarse = nullptr;
  G2P_CUDA(pool_alloc(&d_stats, 2 * sizeof(double), st));
  G2P_CUDA(pool_alloc(&d_coarse, n_coarse, st));
  const double init[2] = {DBL_MAX, 0.0};
  G2P_CUDA(cudaMemcpyAsync(d_stats, init, sizeof(init), cudaMemcpyHostToDevice, st));
  G2P_CUDA(cudaMemsetAsync(d_coarse, 0, n_coarse, st));
  stats_kernel<<<grid_for(n, 256), 256, 0, st>>>(ix->xyz, n, Gc, d_stats, d_coarse);
  G2P_LAUNCHED();
  double stats[2];
  std::vector<uint8_t> coarse(n_coarse);
  G2P_CUDA(cudaMemcpyAsync(stats, d_stats, sizeof(stats), cudaMemcpyDeviceToHost, st));
  G2P_CUDA(cudaMemcpyAsync(coarse.data(), d_coarse, n_coarse, cudaMemcpyDeviceToHost, st));
  G2P_CUDA(cudaStreamSynchronize(st));
  const double t_stats = now_ms();
  pool_free(d_stats, 2 * sizeof(double));
  pool_free(d_coarse, n_coarse);
  const double rmin = stats[0], rmax = stats[1];
  if (!(rmax > 0.0) || !std::isfinite(rmax) || !(rmin > 0.0))
    return set_error(G2P_ERR_INVALID, "source coordinates are not finite / not on a sphere (|p| in [%g, %g])", rmin, rmax);
  if ((rmax - rmin) > 1e-9 * rmax)
    return set_error(G2P_ERR_UNSUPPORTED,
                     "source points must lie on a sphere centred at the origin (|p| in [%.17g, %.17g])", rmin, rmax);
  ix->radius = 0.5 * (rmin + rmax);
  int64_t n_occ = 0;
  for (uint8_t c : coarse) n_occ += c;
  if (n_occ < 1) n_occ = 1;
  const double covered_sr = (double)n_occ * 4.0 * kPi / (double)n_coarse;
  ix->mean_spacing = ix->radius * sqrt(covered_sr / (double)n);
  // 2. grid resolution: about `occ` points per occupied cell
  const double occ = env_double("G2P_CELL_OCC", 1.5);  // measured flat between 0.7 and 1.5 (6.1-6.2 ms), 6.4 ms at 2, 6.5 ms at 3
  double g = (double)Gc * sqrt((double)n / (occ * (double)n_occ));
  int G = (int)floor(g);
  G = G < 2 ? 2 : (G > kMaxG ? kMaxG : G);
  ix->G = G;
  ix->n_cells = 6LL * G * G;
  // 3. counting sort by cell key
  uint32_t *d_keys = nullptr, *d_counts = nullptr, *d_cursor = nullptr, *d_bsums = nullptr;
  const int64_t nc = ix->n_cells;
  const int nb = (int)ceil_div(nc, kScanBlock);
  G2P_CUDA(pool_alloc(&d_keys, (size_t)n * 4, st));
  G2P_CUDA(pool_alloc(&d_counts, (size_t)nc * 4, st));
  G2P_CUDA(pool_alloc(&d_cursor, (size_t)nc * 4, st));
  G2P_CUDA(pool_alloc(&d_bsums, (size_t)nb * 4, st));
  G2P_CUDA(pool_alloc(&ix->cell_start, (size_t)(nc + 1) * 4, st));
  G2P_CUDA(pool_alloc(&ix->pts, (size_t)n * sizeof(Pt), st));
  if (timing) cudaStreamSynchronize(st);
  const double t_alloc = now_ms();
  G2P_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)nc * 4, st));
  key_hist_kernel<<<grid_for(n, 256), 256, 0, st>>>(ix->xyz, n, G, d_keys, d_counts);
  G2P_LAUNCHED();
  scan_blocks_kernel<<<nb, kScanThreads, 0, st>>>(d_counts, ix->cell_start, nc, d_bsums);
  G2P_LAUNCHED();
  scan_sums_kernel<<<1, 1024, 0, st>>>(d_bsums, nb);
  G2P_LAUNCHED();
  scan_add_kernel<<<nb, kScanThreads, 0, st>>>(ix->cell_start, nc, d_bsums, d_cursor, (uint32_t)n);
  G2P_LAUNCHED();
  scatter_kernel<<<grid_for(n, 256), 256, 0, st>>>(ix->xyz, d_keys, n, d_cursor, ix->pts);
  G2P_LAUNCHED();
  // 4. occupied blocks of 16 x 16 cells with their bounding caps: the coarse level of the far search
  const int Gb = (G + kFarB - 1) / kFarB;
  const int max_blocks = 6 * Gb * Gb;
  int* d_nblocks = nullptr;
  G2P_CUDA(pool_alloc(&d_nblocks, sizeof(int), st));
  G2P_CUDA(pool_alloc(&ix->blocks, (size_t)max_blocks * sizeof(FarBlock), st));
  G2P_CUDA(cudaMemsetAsync(d_nblocks, 0, sizeof(int), st));
  far_blocks_kernel<<<(max_blocks + 127) / 128, 128, 0, st>>>(ix->cell_start, G, Gb, ix->blocks, d_nblocks);
  G2P_LAUNCHED();
  G2P_CUDA(cudaMemcpyAsync(&ix->n_blocks, d_nblocks, sizeof(int), cudaMemcpyDeviceToHost, st));
  G2P_CUDA(cudaStreamSynchronize(st));
  pool_free(d_nblocks, sizeof(int));
  ix->blocks_bytes = (size_t)max_blocks * sizeof(FarBlock);
  if (timing)
    fprintf(stderr, "[g2p] index n=%lld: wait-for-xyz %.2f ms, stats+readback %.2f ms, allocations %.2f ms, sort kernels %.2f ms\n",
            (long long)n, t_enter - t0, t_stats - t_enter, t_alloc - t_stats, now_ms() - t_alloc);
  pool_free(d_keys, (size_t)n * 4);
  pool_free(d_counts, (size_t)nc * 4);
  pool_free(d_cursor, (size_t)nc * 4);
  pool_free(d_bsums, (size_t)nb * 4);
  ix->bytes = (int64_t)n * 24 + (int64_t)n * sizeof(Pt) + (nc + 1) * 4;
  return G2P_OK;
}

// ------------------------------------------------------------------ K3  k-NN search
struct IndexView {
  const Pt* pts;
  const uint32_t* cell_start;
  int G;
  int64_t n;
  double radius, h;
  const FarBlock* blocks;   // occupied 16 x 16 cell blocks with bounding caps
  int n_blocks;
};

struct KnnArgs {
  IndexView ix;
  const void* tlon;
  const void* tlat;
  const double* txyz;
  int64_t m;
  int64_t os, oo;            // idx_out / coef_out: row stride and column offset (a block of a larger table)
  Rot rot;
  double xyz_radius;  // radius used by to_3d for the targets (= grid_details['radius'])
  int k, mode;
  double mub;
  int32_t* idx_out;
  double* coef_out;
  uint8_t* flags_out;
  double* xyz_out;
  double* dmax2_out;         // SELF mode: max over queries of d2[1]
  int64_t q0;                // SELF mode: first source record of this call (sharded self query)
  unsigned long long* stats; // [0] extra passes, [1] candidates examined (debug counters; nullable)
  double r0_scale;           // scales the initial search radius (experiments: G2P_KNN_R0)
  double stop2;              // INVDIST: (min_upper_bound * (1 + 1e-6))^2, rows whose nearest point is farther end early; else +inf
  uint32_t* far_rows;        // queries whose cap outgrew the cell search (targets far from a regional source): listed here
  unsigned int* far_count;   // ... and answered by knn_far_kernel; null: exhaustive scan in place
};

__device__ __forceinline__ bool lex_lt(double da, int ia, double db, int ib) {
  return (da < db) || (da == db && ia < ib);
}

// sorted best-KK list plus `nxt`, the smallest squared distance among everything examined that is NOT in the
// list: that one value is all the tie / near-tie flags need from the (k+1)-th neighbour, and keeping KK = k
// slots instead of k+1 saves a quarter of the insertion work for k = 4
template <int KK>
struct TopK {
  double d[KK];
  int id[KK];
  double nxt;
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int s = 0; s < KK; ++s) {
      d[s] = __longlong_as_double(0x7ff0000000000000LL);
      id[s] = INT_MAX;
    }
    nxt = __longlong_as_double(0x7ff0000000000000LL);
  }
  // LEX: order by (distance, id) - the exact rule.  !LEX: order by distance alone (equal distances keep their scan
  // order): a third fewer instructions per inserted candidate; the caller detects exact ties in the result (`has_tie`)
  // and repeats the query with LEX, so the output is the same
  template <bool LEX>
  __device__ __forceinline__ void insert(double nd, int ni) {
    if (LEX ? !lex_lt(nd, ni, d[KK - 1], id[KK - 1]) : !(nd < d[KK - 1])) {
      nxt = fmin(nxt, nd);
      return;
    }
    nxt = fmin(nxt, d[KK - 1]);  // the evicted one (inf while the list is not full)
    d[KK - 1] = nd;
    id[KK - 1] = ni;
#pragma unroll
    for (int s = KK - 1; s > 0; --s) {
      if (LEX ? lex_lt(d[s], id[s], d[s - 1], id[s - 1]) : (d[s] < d[s - 1])) {
        double td = d[s]; d[s] = d[s - 1]; d[s - 1] = td;
        int ti = id[s]; id[s] = id[s - 1]; id[s - 1] = ti;
      }
    }
  }
  // two equal finite distances among the list and the best one behind it
  __device__ __forceinline__ bool has_tie() const {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    bool t = nxt < inf && nxt == d[KK - 1];
#pragma unroll
    for (int s = 0; s + 1 < KK; ++s) t = t || (d[s + 1] < inf && d[s] == d[s + 1]);
    return t;
  }
  __device__ __forceinline__ void pop_front() {
#pragma unroll
    for (int s = 0; s < KK - 1; ++s) {
      d[s] = d[s + 1];
      id[s] = id[s + 1];
    }
    d[KK - 1] = __longlong_as_double(0x7ff0000000000000LL);
    id[KK - 1] = INT_MAX;
  }
};

template <int KK, int LANES, bool LEX>
__device__ __forceinline__ void scan_range(const Pt* __restrict__ pts, uint32_t p0, uint32_t p1, int lane, double qx,
                                           double qy, double qz, TopK<KK>& top, unsigned& examined) {
  for (uint32_t p = p0 + lane; p < p1; p += LANES) {
    double px, py, pz;
    int id;
    load_pt(pts + p, px, py, pz, id);
    const double dx = __dsub_rn(qx, px), dy = __dsub_rn(qy, py), dz = __dsub_rn(qz, pz);
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    top.template insert<LEX>(d2, id);
    ++examined;
  }
}

// (alpha or beta) range on one face axis of the cap with sin(gamma) = s around unit q; returns false if empty.
// L = sqrt(a^2 + c^2) is the length of q's projection on the plane spanned by the axis pair.
// asin for the small arguments of a search cap (its radius is a few grid spacings): the series x + x^3/6 + 3x^5/40 is
// within 1.3e-10 of asin(x) for x < 0.06, far inside the 4e-6 slack the callers add; asinf() beyond
__device__ __forceinline__ float asin_small(float x) {
  if (x < 0.06f) {
    const float x2 = x * x;
    return x * (1.0f + x2 * (0.16666667f + x2 * 0.075f));
  }
  return asinf(x);
}

__device__ __forceinline__ bool axis_range(float a, float c, float s, int G, int& lo, int& hi) {
  const float L = sqrtf(a * a + c * c);
  float amin = -(float)kQuarterPi, amax = (float)kQuarterPi;
  if (c > 0.0f && s < 0.70f * L) {
    const float th = atanf(a / c);
    const float de = asin_small(s / L) * 1.00001f + 4e-6f;  // slack >> float error of th/de, << a cell
    amin = fmaxf(amin, th - de);
    amax = fminf(amax, th + de);
    if (amin > amax) return false;
  }
  const double scale = (double)G / (kPi / 2.0);
  int l = (int)floor(((double)amin + kQuarterPi) * scale);
  int h = (int)floor(((double)amax + kQuarterPi) * scale);
  lo = l < 0 ? 0 : (l >= G ? G - 1 : l);
  hi = h < 0 ? 0 : (h >= G ? G - 1 : h);
  return true;
}

// Examine every source point within Euclidean distance D of q (and possibly more).
template <int KK, int LANES, bool LEX>
__device__ __forceinline__ void scan_cap(const IndexView& ix, double qx, double qy, double qz, double qnorm, double D,
                                         int lane, TopK<KK>& top, unsigned& examined, bool& whole, bool far_ok, bool& far) {
  // |p-q|^2 = R^2 + rho^2 - 2 R rho cos(g)  =>  cos(g) >= c
  const double R = ix.radius, rho = qnorm;
  const double c = (R * R + rho * rho - D * D) / (2.0 * R * rho);
  if (c > 1.0 + 1e-9) return;  // the sphere is farther than D from q
  if (!(c > 0.82) || !(rho > 0.0)) {
    if (far_ok) {   // the coarse-to-fine pass takes this query (knn_far_kernel)
      far = true;
      whole = true;
      return;
    }
    // cap wider than ~35 degrees (or a degenerate query): exhaustive scan
    scan_range<KK, LANES, LEX>(ix.pts, 0u, (uint32_t)ix.n, lane, qx, qy, qz, top, examined);
    whole = true;
    return;
  }
  const double cc = c > 1.0 ? 1.0 : c;
  const float s = (float)(sqrt(1.0 - cc * cc) * (1.0 + 1e-6)) + 1e-7f;
  const float gam = asin_small(fminf(s, 1.0f));
  const double inv = 1.0 / rho;
  const double ux = qx * inv, uy = qy * inv, uz = qz * inv;
  const int G = ix.G;
#pragma unroll 1
  for (int f = 0; f < 6; ++f) {
    double a, b, cf;
    face_local(f, ux, uy, uz, a, b, cf);
    // the face lies within 54.74 deg of its axis: unreachable if angle(q, axis) > gamma + 54.74 deg
    if ((float)cf < cosf(fminf(gam + 0.9554f, 3.1415926f)) - 1e-5f) continue;
    int alo, ahi, blo, bhi;
    if (LANES == 2) {
      // the two lanes of a query hold the same q: each works out one axis of the rectangle and they swap the results
      const unsigned pair = 3u << ((threadIdx.x & 31) & ~1);
      int lo, hi;
      if (!axis_range((float)(lane ? b : a), (float)cf, s, G, lo, hi)) {
        lo = 1;
        hi = 0;
      }
      const int olo = __shfl_xor_sync(pair, lo, 1), ohi = __shfl_xor_sync(pair, hi, 1);
      alo = lane ? olo : lo;
      ahi = lane ? ohi : hi;
      blo = lane ? lo : olo;
      bhi = lane ? hi : ohi;
      if (alo > ahi || blo > bhi) continue;
    } else {
      if (!axis_range((float)a, (float)cf, s, G, alo, ahi)) continue;
      if (!axis_range((float)b, (float)cf, s, G, blo, bhi)) continue;
    }
    for (int ib = blo; ib <= bhi; ++ib) {
      const int64_t row = ((int64_t)f * G + ib) * G;
      const uint32_t p0 = __ldg(ix.cell_start + row + alo);
      const uint32_t p1 = __ldg(ix.cell_start + row + ahi + 1);
      scan_range<KK, LANES, LEX>(ix.pts, p0, p1, lane, qx, qy, qz, top, examined);
    }
  }
}

template <int KK, int LANES>
__device__ __forceinline__ void merge_group(TopK<KK>& top, unsigned gmask) {
  if (LANES == 1) return;
  TopK<KK> res;
#pragma unroll
  for (int s = 0; s < KK; ++s) {
    double bd = top.d[0];
    int bi = top.id[0];
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(gmask, bd, o);
      const int oi = __shfl_xor_sync(gmask, bi, o);
      if (lex_lt(od, oi, bd, bi)) {
        bd = od;
        bi = oi;
      }
    }
    res.d[s] = bd;
    res.id[s] = bi;
    if (top.id[0] == bi && bi != INT_MAX) top.pop_front();
  }
  // best of what is left anywhere in the group: the lanes' unselected heads and their own `nxt`
  double nx = fmin(top.nxt, top.d[0]);
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) nx = fmin(nx, __shfl_xor_sync(gmask, nx, o));
  res.nxt = nx;
  top = res;
}

// initial search radius in units of the mean spacing h, by k
__device__ __forceinline__ double initial_radius_factor(int k) {
  // (measured on O1280 -> 0.05 deg, k = 4: 1.44 h -> 8.00 ms per build, 1.6 h -> 8.24 ms, 1.28 h -> 8.44 ms, 1.76 h -> 8.56 ms:
  // a smaller cap scans fewer cells but more queries need a second pass)
  return 0.9 * (k <= 1 ? 0.85 : (k == 2 ? 1.25 : (k <= 4 ? 1.6 : 0.62 * sqrt((double)k) + 0.5)));
}

template <int KK, int LANES, bool LEX>
__device__ __forceinline__ void knn_search(const IndexView& ix, int k, double qx, double qy, double qz, int lane,
                                           unsigned gmask, TopK<KK>& top, unsigned& passes, unsigned& examined,
                                           double r0_scale, double stop2, bool far_ok, bool& far) {
  const double qnorm = sqrt(qx * qx + qy * qy + qz * qz);
  double D = ix.h * initial_radius_factor(k) * r0_scale;
  const double Dmax = 2.0 * (ix.radius + qnorm) + ix.radius;
  passes = 0;
  examined = 0;
  for (;;) {
    top.reset();
    bool whole = false;
    scan_cap<KK, LANES, LEX>(ix, qx, qy, qz, qnorm, D, lane, top, examined, whole, far_ok, far);
    if (far) return;
    merge_group<KK, LANES>(top, gmask);
    ++passes;
    const double dk2 = top.d[k - 1];
    if (whole || dk2 <= D * D) return;
    // invdist only (stop2 = min_upper_bound^2 with a safety margin, else +inf): the row is "outside" as soon as the nearest
    // source point is known to lie beyond min_upper_bound - either it was found inside the cap (then it IS the
    // nearest) or the cap is empty and already wider than the bound.  Far targets of a regional source end here
    // instead of widening the cap up to the exhaustive scan; the epilogue only looks at d0 for such rows.
    if (top.d[0] <= D * D ? top.d[0] > stop2 : D * D > stop2) return;
    if (dk2 < __longlong_as_double(0x7ff0000000000000LL)) {
      D = sqrt(dk2) * (1.0 + 1e-9);  // k candidates known: the cap through the k-th one is final
    } else {
      // fewer than k points within eight times the initial radius: the query lies outside the source domain (or in a
      // hole of it).  Wider caps walk thousands of empty cell rows - the coarse-to-fine far pass answers such a query
      // at a fixed, small cost whatever its distance
      if (far_ok && passes >= 4) {
        far = true;
        return;
      }
      D *= 2.0;
      if (D > Dmax) D = Dmax * 4.0;  // forces the exhaustive branch
    }
  }
}

// K5 epilogue arithmetic (also used by weights_kernel).  d[] are float64 distances.
template <bool F32>
__device__ __forceinline__ void invdist_weights(const double* d, int k, double mub, double* w, bool& inside) {
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  if (F32) {
    const float d0 = (float)d[0];
    if (d0 <= 1e-10f) {
      for (int j = 0; j < k; ++j) w[j] = (j == 0) ? 1.0 : 0.0;
      inside = true;
    } else if ((double)d0 <= mub) {
      float wf[16];
      float sum = 0.0f;
      for (int j = 0; j < k; ++j) {
        const float dj = (float)d[j];
        wf[j] = __fdiv_rn(1.0f, __fmul_rn(dj, dj));
        sum = (j == 0) ? wf[0] : __fadd_rn(sum, wf[j]);
      }
      for (int j = 0; j < k; ++j) w[j] = (double)__fdiv_rn(wf[j], sum);
      inside = true;
    } else {
      for (int j = 0; j < k; ++j) w[j] = nan;
      inside = false;
    }
  } else {
    if (d[0] <= 1e-10) {
      for (int j = 0; j < k; ++j) w[j] = (j == 0) ? 1.0 : 0.0;
      inside = true;
    } else if (d[0] <= mub) {
      double wd[16];
      double sum = 0.0;
      for (int j = 0; j < k; ++j) {
        wd[j] = __ddiv_rn(1.0, __dmul_rn(d[j], d[j]));
        sum = (j == 0) ? wd[0] : __dadd_rn(sum, wd[j]);
      }
      for (int j = 0; j < k; ++j) w[j] = __ddiv_rn(wd[j], sum);
      inside = true;
    } else {
      for (int j = 0; j < k; ++j) w[j] = nan;
      inside = false;
    }
  }
}

enum { QUERY_LATLON_F64 = 0, QUERY_LATLON_F32 = 1, QUERY_XYZ = 2, QUERY_SELF = 3 };

// one row of the output from a finished (merged) list: flags, distances, the mode's indexes / coefficients.  Called by one
// lane per query.
template <int KK, bool F32>
__device__ __forceinline__ void emit_row(const KnnArgs& a, int64_t q, int k, const TopK<KK>& top, double qx, double qy,
                                         double qz) {
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
    if (a.xyz_out) {
    a.xyz_out[3 * q] = qx; a.xyz_out[3 * q + 1] = qy; a.xyz_out[3 * q + 2] = qz;
  }
  uint8_t flag = 0;
  double dist[KK];
#pragma unroll
  for (int s = 0; s < KK; ++s) {
    if (s < k) {
      double d = __dsqrt_rn(top.d[s]);
      if (F32) d = (double)(float)d;
      dist[s] = d;
      if (top.id[s] == INT_MAX) flag |= G2P_FLAG_SHORT;
      // the next-best squared distance after slot s: the following slot, or `nxt` behind the last one
      const double after = (s + 1 < KK) ? top.d[(s + 1 < KK) ? s + 1 : s] : top.nxt;
      if (after < inf) {
        if (top.d[s] == after) flag |= G2P_FLAG_TIE;
        else if (after - top.d[s] <= after * 5.7e-14) flag |= G2P_FLAG_NEAR_TIE;
      }
    }
  }
  if (a.flags_out) a.flags_out[q] = flag;
  const int32_t nsrc = (int32_t)a.ix.n;
  if (a.mode == G2P_MODE_RAW) {
#pragma unroll
    for (int s = 0; s < KK; ++s)
      if (s < k) {
        a.idx_out[(int64_t)s * a.os + a.oo + q] = (top.id[s] == INT_MAX) ? nsrc : top.id[s];
        a.coef_out[(int64_t)s * a.os + a.oo + q] = dist[s];
      }
  } else if (a.mode == G2P_MODE_NEAREST) {
    // _build_nn: idx = ix if dist <= mub else N ; coeffs = distance for every row
    a.idx_out[a.oo + q] = (dist[0] <= a.mub && top.id[0] != INT_MAX) ? top.id[0] : nsrc;
    a.coef_out[a.oo + q] = dist[0];
  } else {
    double w[KK];
    bool inside;
    invdist_weights<F32>(dist, k, a.mub, w, inside);
#pragma unroll
    for (int s = 0; s < KK; ++s)
      if (s < k) {
        a.idx_out[(int64_t)s * a.os + a.oo + q] = (inside && top.id[s] != INT_MAX) ? top.id[s] : nsrc;
        a.coef_out[(int64_t)s * a.os + a.oo + q] = w[s];
      }
  }
  }



template <int KK, int LANES, int QSRC>
__global__ void __launch_bounds__(256, KK <= 2 ? 5 : (KK <= 4 ? 4 : (KK < 11 ? 3 : 1))) knn_kernel(const KnnArgs a) {
  constexpr bool F32 = (QSRC == QUERY_LATLON_F32);
  const int lane = threadIdx.x % LANES;
  const unsigned gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << ((threadIdx.x & 31) / LANES * LANES));
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  const int64_t n_groups = (int64_t)gridDim.x * blockDim.x / LANES;
  const int k = a.k;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  double self_max = 0.0;
  unsigned long long extra_passes = 0, cand = 0;
  for (int64_t q = group; q < a.m; q += n_groups) {
    double qx, qy, qz;
    if (QSRC == QUERY_SELF) {
      int self_id;
      load_pt(a.ix.pts + a.q0 + q, qx, qy, qz, self_id);
    } else if (QSRC == QUERY_XYZ) {
      qx = a.txyz[3 * q]; qy = a.txyz[3 * q + 1]; qz = a.txyz[3 * q + 2];
    } else {
      latlon_to_xyz<F32>(a.tlon, a.tlat, q, a.xyz_radius, a.rot, qx, qy, qz, LANES >= 2 ? gmask : 0u);
    }
    TopK<KK> top;
    unsigned passes, examined;
    bool far = false;
    const bool far_ok = QSRC != QUERY_SELF && a.far_rows != nullptr;
    knn_search<KK, LANES, false>(a.ix, k, qx, qy, qz, lane, gmask, top, passes, examined, a.r0_scale, a.stop2, far_ok, far);
    if (far) {   // every lane of the query takes this branch
      if (lane == 0) a.far_rows[atomicAdd(a.far_count, 1u)] = (uint32_t)q;
      continue;
    }
    if (QSRC != QUERY_SELF && top.has_tie()) {
      // exact float64 ties (54 of 21.6 M rows on O1280 -> 0.05 deg): once more with the (distance, id) order, which
      // decides the order inside the list and who stays at its end.  All lanes of a query see the same merged list.
      knn_search<KK, LANES, true>(a.ix, k, qx, qy, qz, lane, gmask, top, passes, examined, a.r0_scale, a.stop2, false, far);
    }
    extra_passes += passes - 1;
    cand += examined;
    if (QSRC == QUERY_SELF) {
      // np.max over both columns of the k=2 self query: d[:,0] is 0 (the point itself)
      const double d2 = top.d[1];
      if (lane == 0 && d2 > self_max) self_max = d2;
      continue;
    }
    if (LANES == 2 && KK == 4 && k == 4 && a.mode == G2P_MODE_INVDIST) {
      // invdist epilogue on BOTH lanes of the query (each holds the same merged list): lane l takes the square roots,
      // reciprocals and divisions of neighbours 2l and 2l+1; only the running sum ((w0 + w1) + w2) + w3 crosses lanes
      const double e2a = lane ? top.d[2] : top.d[0], e2b = lane ? top.d[3] : top.d[1];
      const int ida = lane ? top.id[2] : top.id[0], idb = lane ? top.id[3] : top.id[1];
      double da = __dsqrt_rn(e2a), db = __dsqrt_rn(e2b);
      if (F32) {
        da = (double)(float)da;
        db = (double)(float)db;
      }
      const int lead = (threadIdx.x & 31) & ~1;                 // lane 0 of this pair within the warp
      const double d0 = __shfl_sync(gmask, da, lead);
      const double nan = __longlong_as_double(0x7ff8000000000000LL);
      double wa, wb;
      bool inside;
      if (F32 ? ((float)d0 <= 1e-10f) : (d0 <= 1e-10)) {        // the target coincides with a source point
        wa = lane ? 0.0 : 1.0;
        wb = 0.0;
        inside = true;
      } else if (d0 <= a.mub) {   // (both lanes of a pair see the same d0: they take the same branch, shuffles included)
        if (F32) {
          const float fa = (float)da, fb = (float)db;
          const float ra = __fdiv_rn(1.0f, __fmul_rn(fa, fa)), rb = __fdiv_rn(1.0f, __fmul_rn(fb, fb));
          const float p01 = __shfl_sync(gmask, __fadd_rn(ra, rb), lead);      // w0 + w1 (from lane 0)
          const float sum = __shfl_sync(gmask, __fadd_rn(__fadd_rn(p01, ra), rb), lead + 1);   // (+ w2) + w3 (on lane 1)
          wa = (double)__fdiv_rn(ra, sum);
          wb = (double)__fdiv_rn(rb, sum);
        } else {
          const double ra = __ddiv_rn(1.0, __dmul_rn(da, da)), rb = __ddiv_rn(1.0, __dmul_rn(db, db));
          const double p01 = __shfl_sync(gmask, __dadd_rn(ra, rb), lead);
          const double sum = __shfl_sync(gmask, __dadd_rn(__dadd_rn(p01, ra), rb), lead + 1);
          wa = __ddiv_rn(ra, sum);
          wb = __ddiv_rn(rb, sum);
        }
        inside = true;
      } else {
        wa = wb = nan;
        inside = false;
      }
      const int32_t nsrc2 = (int32_t)a.ix.n;
      const int64_t sa = (int64_t)(2 * lane) * a.os + a.oo + q;
      a.idx_out[sa] = (inside && ida != INT_MAX) ? ida : nsrc2;
      a.coef_out[sa] = wa;
      a.idx_out[sa + a.os] = (inside && idb != INT_MAX) ? idb : nsrc2;
      a.coef_out[sa + a.os] = wb;
      if (lane == 0) {
        if (a.xyz_out) {
          a.xyz_out[3 * q] = qx; a.xyz_out[3 * q + 1] = qy; a.xyz_out[3 * q + 2] = qz;
        }
        if (a.flags_out) {
          uint8_t flag = 0;
#pragma unroll
          for (int s = 0; s < KK; ++s) {
            if (top.id[s] == INT_MAX) flag |= G2P_FLAG_SHORT;
            const double after = (s + 1 < KK) ? top.d[(s + 1 < KK) ? s + 1 : s] : top.nxt;
            if (after < inf) {
              if (top.d[s] == after) flag |= G2P_FLAG_TIE;
              else if (after - top.d[s] <= after * 5.7e-14) flag |= G2P_FLAG_NEAR_TIE;
            }
          }
          a.flags_out[q] = flag;
        }
      }
      continue;
    }
    if (lane != 0) continue;
    emit_row<KK, F32>(a, q, k, top, qx, qy, qz);
  }
  if (QSRC == QUERY_SELF) {
    for (int o = 16; o > 0; o >>= 1) self_max = fmax(self_max, __shfl_xor_sync(0xffffffffu, self_max, o));
    if ((threadIdx.x & 31) == 0) atomic_max_pos_double(a.dmax2_out, self_max);
  }
  if (a.stats && lane == 0) {
    atomicAdd(&a.stats[0], extra_passes);
    atomicAdd(&a.stats[1], cand);
  }
}

// ---- the far pass: queries whose cap outgrew the cell search (targets far from a regional source: `nearest` stores the
// true distance of every row, App. C.4, and RAW / ADW need the neighbours).  One warp per listed query, coarse to fine:
// the occupied 16 x 16 cell blocks carry a bounding cap, so |q - p|^2 >= R^2 + rho^2 - 2 R rho cos(max(0, theta - r))
// for every point p of a block at angle theta from q.  The block with the smallest upper bound is scanned first; then
// every block whose lower bound does not exceed the current k-th best (the smallest k-th best of any lane bounds the
// warp's) is scanned by all 32 lanes.  Distances and ordering are the main kernel's (float64, no FMA, (d2, id) order), so
// the rows are the ones the exhaustive scan produced - at ~1e3 instead of n distance evaluations per query.
template <int KK>
__device__ __forceinline__ void far_scan_block(const IndexView& ix, const FarBlock& fb, int lane, double qx, double qy,
                                               double qz, TopK<KK>& top) {
  const int G = ix.G;
  const int ia1 = min(fb.ia0 + kFarB, G), ib1 = min(fb.ib0 + kFarB, G);
  // lane r < kFarB holds the record range of cell row ib0 + r; a prefix sum over those lanes numbers the block's points
  // 0 .. total-1, and the 32 lanes walk that numbering together (one round of dependent loads instead of one per row)
  uint32_t p0 = 0, p1 = 0;
  if (lane < kFarB && fb.ib0 + lane < ib1) {
    const int64_t row = ((int64_t)fb.f * G + fb.ib0 + lane) * G;
    p0 = __ldg(ix.cell_start + row + fb.ia0);
    p1 = __ldg(ix.cell_start + row + ia1);
  }
  const int cnt = (int)(p1 - p0);
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < kFarB; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const int total = __shfl_sync(0xffffffffu, incl, kFarB - 1);
  const int excl = incl - cnt;
  for (int base = 0; base < total; base += 32) {
    const int t = base + lane;
    int j = 0;   // the last row whose first point number is <= t
#pragma unroll
    for (int step = kFarB / 2; step >= 1; step >>= 1) {
      const int e = __shfl_sync(0xffffffffu, excl, j + step);
      if (t >= e) j += step;
    }
    const uint32_t first = __shfl_sync(0xffffffffu, p0, j);
    const int e0 = __shfl_sync(0xffffffffu, excl, j);
    if (t < total) {
      double px, py, pz;
      int id;
      load_pt(ix.pts + first + (uint32_t)(t - e0), px, py, pz, id);
      const double dx = __dsub_rn(qx, px), dy = __dsub_rn(qy, py), dz = __dsub_rn(qz, pz);
      const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      top.template insert<true>(d2, id);
    }
  }
}

template <int KK, int QSRC>
__global__ void __launch_bounds__(256) knn_far_kernel(const KnnArgs a) {
  constexpr bool F32 = (QSRC == QUERY_LATLON_F32);
  const int lane = threadIdx.x & 31;
  const unsigned n_far = *a.far_count;
  const int warps = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  const int k = a.k;
  for (unsigned w = (unsigned)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); w < n_far; w += warps) {
    const int64_t q = a.far_rows[w];
    double qx, qy, qz;
    if (QSRC == QUERY_XYZ) {
      qx = a.txyz[3 * q]; qy = a.txyz[3 * q + 1]; qz = a.txyz[3 * q + 2];
    } else {
      latlon_to_xyz<F32>(a.tlon, a.tlat, q, a.xyz_radius, a.rot, qx, qy, qz, 0u);   // the main kernel's bits
    }
    const double rho = sqrt(qx * qx + qy * qy + qz * qz), R = a.ix.radius;
    const double inv = rho > 0.0 ? 1.0 / rho : 0.0;
    const float ux = (float)(qx * inv), uy = (float)(qy * inv), uz = (float)(qz * inv);
    const double base = R * R + rho * rho, two = 2.0 * R * rho;
    // bounds of a block: cos of (theta -+ r), clamped; 1e-5 absorbs the float32 rounding of the dot product
    auto bounds = [&](const FarBlock& fb, double& lb2, double& ub2) {
      float ct = fminf(1.0f, fmaxf(-1.0f, ux * fb.cx + uy * fb.cy + uz * fb.cz));
      const float st = sqrtf(fmaxf(0.0f, 1.0f - ct * ct));
      const float c_near = (ct >= fb.cos_r) ? 1.0f : fminf(1.0f, ct * fb.cos_r + st * fb.sin_r + 1e-5f);   // cos(theta - r)
      const float c_far = (ct <= -fb.cos_r) ? -1.0f : fmaxf(-1.0f, ct * fb.cos_r - st * fb.sin_r - 1e-5f);  // cos(theta + r)
      lb2 = fmax(0.0, base - two * (double)c_near) * (1.0 - 1e-6);
      ub2 = (base - two * (double)c_far) * (1.0 + 1e-6);
    };
    TopK<KK> top;
    top.reset();
    // seed: the block with the smallest upper bound
    double best_ub = inf;
    int best_b = -1;
    for (int b = lane; b < a.ix.n_blocks; b += 32) {
      double lb2, ub2;
      bounds(a.ix.blocks[b], lb2, ub2);
      if (ub2 < best_ub) {
        best_ub = ub2;
        best_b = b;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ou = __shfl_xor_sync(0xffffffffu, best_ub, o);
      const int ob = __shfl_xor_sync(0xffffffffu, best_b, o);
      if (ou < best_ub || (ou == best_ub && ob >= 0 && (best_b < 0 || ob < best_b))) {
        best_ub = ou;
        best_b = ob;
      }
    }
    if (best_b >= 0) far_scan_block<KK>(a.ix, a.ix.blocks[best_b], lane, qx, qy, qz, top);
    auto kth_bound = [&]() {   // an upper bound of the warp's k-th best: the smallest k-th best of any lane
      double v = top.d[k - 1];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
      return v;
    };
    double B = kth_bound();
    for (int b0 = 0; b0 < a.ix.n_blocks; b0 += 32) {
      const int b = b0 + lane;
      double lb2 = inf, ub2;
      bool need = b < a.ix.n_blocks && b != best_b;      // (B is +inf while fewer than k points are known)
      if (need) bounds(a.ix.blocks[b], lb2, ub2);
      for (;;) {
        const unsigned mask = __ballot_sync(0xffffffffu, need && lb2 <= B);
        if (!mask) break;
        const int src = __ffs(mask) - 1;
        far_scan_block<KK>(a.ix, a.ix.blocks[b0 + src], lane, qx, qy, qz, top);
        if (lane == src) need = false;
        B = kth_bound();
      }
    }
    merge_group<KK, 32>(top, 0xffffffffu);
    if (lane == 0) emit_row<KK, F32>(a, q, k, top, qx, qy, qz);
  }
}

// standalone K5
template <bool F32>
__global__ void weights_kernel(int mode, int k, int64_t m, int32_t nsrc, double mub, int32_t* __restrict__ idx,
                               double* __restrict__ coef) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < m; q += (int64_t)gridDim.x * blockDim.x) {
    if (mode == G2P_MODE_NEAREST) {
      double d = coef[q];
      if (F32) d = (double)(float)d;
      coef[q] = d;
      if (!(d <= mub)) idx[q] = nsrc;
    } else {
      double d[16], w[16];
      for (int j = 0; j < k; ++j) d[j] = coef[(int64_t)j * m + q];
      bool inside;
      invdist_weights<F32>(d, k, mub, w, inside);
      for (int j = 0; j < k; ++j) {
        coef[(int64_t)j * m + q] = w[j];
        if (!inside) idx[(int64_t)j * m + q] = nsrc;
      }
    }
  }
}

static int knn_lanes(int kk) {
  static int forced = -1;
  if (forced < 0) {
    const int v = (int)env_double("G2P_KNN_LANES", 0);
    forced = (v == 1 || v == 2 || v == 4 || v == 8 || v == 32) ? v : 0;
  }
  if (forced) return forced;
  // threads per query.  Measured on B200, O1280 -> 0.05 deg k=4 (21.6 M queries): 1 lane 17.8 ms, 2 lanes 7.5 ms,
  // 4 lanes 12.8 ms, 8 lanes 23.4 ms - with ~20 candidates per query the shuffle merge of wider groups costs more
  // than the scan it parallelises; for one or two neighbours a lane of its own per query is best (k=1, 5.4 M
  // queries: 0.65 ms against 0.82 ms; the k=2 self query of 6.6 M points: 0.71 ms against 0.87 ms), profiles/r01_build.md
  return kk <= 2 ? 1 : 2;
}

template <int KK, int LANES>
static int launch_knn_q(const KnnArgs& a, int qsrc, cudaStream_t st) {
  const int threads = 256;
  const int64_t groups_per_block = threads / LANES;
  int64_t blocks = ceil_div(a.m, groups_per_block);
  int64_t cap = (int64_t)num_sms() * 32;
  int grid = (int)(blocks < cap ? blocks : cap);
  switch (qsrc) {
    case QUERY_LATLON_F64: knn_kernel<KK, LANES, QUERY_LATLON_F64><<<grid, threads, 0, st>>>(a); break;
    case QUERY_LATLON_F32: knn_kernel<KK, LANES, QUERY_LATLON_F32><<<grid, threads, 0, st>>>(a); break;
    case QUERY_XYZ: knn_kernel<KK, LANES, QUERY_XYZ><<<grid, threads, 0, st>>>(a); break;
    default: knn_kernel<KK, LANES, QUERY_SELF><<<grid, threads, 0, st>>>(a); break;
  }
  G2P_LAUNCHED();
  return G2P_OK;
}

template <int KK>
static int launch_knn_l(const KnnArgs& a, int qsrc, cudaStream_t st) {
  switch (knn_lanes(KK)) {
    case 1: return launch_knn_q<KK, 1>(a, qsrc, st);
    case 2: return launch_knn_q<KK, 2>(a, qsrc, st);
    case 4: return launch_knn_q<KK, 4>(a, qsrc, st);
    case 32: return launch_knn_q<KK, 32>(a, qsrc, st);
    case 8: return launch_knn_q<KK, 8>(a, qsrc, st);
    default: return launch_knn_q<KK, 2>(a, qsrc, st);
  }
}

static int launch_knn(const KnnArgs& a, int qsrc, cudaStream_t st) {
  const int k = a.k;
  if (k <= 1) return launch_knn_l<1>(a, qsrc, st);
  if (k <= 2) return launch_knn_l<2>(a, qsrc, st);
  if (k <= 4) return launch_knn_l<4>(a, qsrc, st);
  if (k <= 11) return launch_knn_l<11>(a, qsrc, st);
  return launch_knn_l<16>(a, qsrc, st);
}

template <int KK>
static int launch_far_k(const KnnArgs& a, int qsrc, cudaStream_t st) {
  const int grid = num_sms() * 4;   // reads the number of listed queries on the device: no host round trip
  switch (qsrc) {
    case QUERY_LATLON_F64: knn_far_kernel<KK, QUERY_LATLON_F64><<<grid, 256, 0, st>>>(a); break;
    case QUERY_LATLON_F32: knn_far_kernel<KK, QUERY_LATLON_F32><<<grid, 256, 0, st>>>(a); break;
    default: knn_far_kernel<KK, QUERY_XYZ><<<grid, 256, 0, st>>>(a); break;
  }
  G2P_LAUNCHED();
  return G2P_OK;
}

static int launch_far(const KnnArgs& a, int qsrc, cudaStream_t st) {
  if (a.k <= 1) return launch_far_k<1>(a, qsrc, st);
  if (a.k <= 4) return launch_far_k<4>(a, qsrc, st);
  return launch_far_k<16>(a, qsrc, st);
}

static IndexView view_of(const g2p_index* ix) {
  IndexView v;
  v.pts = ix->pts;
  v.cell_start = ix->cell_start;
  v.G = ix->G;
  v.n = ix->n;
  v.radius = ix->radius;
  v.h = ix->mean_spacing;
  v.blocks = ix->blocks;
  v.n_blocks = ix->n_blocks;
  return v;
}

// ------------------------------------------------------------------ table compaction (touched source points)
__global__ void mark_touched_kernel(const int32_t* __restrict__ idx, int64_t total, int64_t n, uint32_t* __restrict__ used) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t v = idx[i];
    if (v >= 0 && v < n) used[v] = 1u;  // benign race: every writer stores the same value
  }
}
__global__ void list_touched_kernel(const uint32_t* __restrict__ used, const uint32_t* __restrict__ pos, int64_t n,
                                    int32_t* __restrict__ touched) {
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x)
    if (used[v]) touched[pos[v]] = (int32_t)v;
}
__global__ void renumber_kernel(const int32_t* __restrict__ idx, int64_t total, int64_t n, const uint32_t* __restrict__ pos,
                                int32_t nc, int32_t* __restrict__ cidx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t v = idx[i];
    cidx[i] = (v >= 0 && v < n) ? (int32_t)pos[v] : nc;
  }
}

}  // namespace g2p

using namespace g2p;

extern "C" {

int g2p_latlon_to_xyz(const void* lon, const void* lat, int64_t n, int dtype, double radius, const g2p_rotation* rot,
                      double* xyz_out, void* stream) {
  if (!lon || !lat || !xyz_out || n < 0) return set_error(G2P_ERR_INVALID, "g2p_latlon_to_xyz: bad arguments");
  if (dtype != G2P_F64 && dtype != G2P_F32) return set_error(G2P_ERR_INVALID, "g2p_latlon_to_xyz: bad dtype");
  if (n == 0) return G2P_OK;
  Rot r = make_rot(rot);
  cudaStream_t st = as_stream(stream);
  if (dtype == G2P_F32)
    xyz_kernel<true><<<grid_for(n, 256), 256, 0, st>>>(lon, lat, n, radius, r, xyz_out);
  else
    xyz_kernel<false><<<grid_for(n, 256), 256, 0, st>>>(lon, lat, n, radius, r, xyz_out);
  G2P_LAUNCHED();
  return G2P_OK;
}

static int index_alloc(int64_t n, g2p_index** out, g2p_index** ixp, cudaStream_t st) {
  if (!out) return set_error(G2P_ERR_INVALID, "g2p_index_create: out is null");
  *out = nullptr;
  if (n < 1 || n > 0x7ffffff0LL) return set_error(G2P_ERR_INVALID, "g2p_index_create: n=%lld out of range", (long long)n);
  g2p_index* ix = new g2p_index();
  ix->n = n;
  G2P_CUDA(cudaGetDevice(&ix->device));
  cudaError_t e = pool_alloc(&ix->xyz, (size_t)n * 24, st);
  if (e != cudaSuccess) {
    delete ix;
    return cuda_fail(e, "cudaMalloc(xyz)", __FILE__, __LINE__);
  }
  *ixp = ix;
  return G2P_OK;
}

int g2p_index_create(const double* lon, const double* lat, int64_t n, double radius, const g2p_rotation* rot,
                     g2p_index** out, void* stream) {
  if (!lon || !lat) return set_error(G2P_ERR_INVALID, "g2p_index_create: null lon/lat");
  g2p_index* ix = nullptr;
  int rc = index_alloc(n, out, &ix, as_stream(stream));
  if (rc) return rc;
  ix->user_radius = radius;
  rc = g2p_latlon_to_xyz(lon, lat, n, G2P_F64, radius, rot, ix->xyz, stream);
  if (rc == G2P_OK) rc = build_cells(ix, as_stream(stream));
  if (rc != G2P_OK) {
    free_index(ix);
    return rc;
  }
  *out = ix;
  return G2P_OK;
}

int g2p_index_create_xyz(const double* xyz, int64_t n, g2p_index** out, void* stream) {
  if (!xyz) return set_error(G2P_ERR_INVALID, "g2p_index_create_xyz: null xyz");
  g2p_index* ix = nullptr;
  int rc = index_alloc(n, out, &ix, as_stream(stream));
  if (rc) return rc;
  cudaError_t e = cudaMemcpyAsync(ix->xyz, xyz, (size_t)n * 24, cudaMemcpyDeviceToDevice, as_stream(stream));
  rc = (e == cudaSuccess) ? build_cells(ix, as_stream(stream)) : cuda_fail(e, "cudaMemcpyAsync(xyz)", __FILE__, __LINE__);
  if (rc != G2P_OK) {
    free_index(ix);
    return rc;
  }
  *out = ix;
  return G2P_OK;
}

int g2p_index_destroy(g2p_index* index) {
  free_index(index);
  return G2P_OK;
}

int g2p_index_get_info(const g2p_index* ix, g2p_index_info* info) {
  if (!ix || !info) return set_error(G2P_ERR_INVALID, "g2p_index_get_info: null argument");
  info->n = ix->n;
  info->cells_per_face_side = ix->G;
  info->n_cells = ix->n_cells;
  info->radius = ix->radius;
  info->mean_spacing = ix->mean_spacing;
  info->bytes = ix->bytes;
  return G2P_OK;
}

int g2p_index_get_xyz(const g2p_index* ix, double* xyz_out, void* stream) {
  if (!ix || !xyz_out) return set_error(G2P_ERR_INVALID, "g2p_index_get_xyz: null argument");
  G2P_CUDA(cudaMemcpyAsync(xyz_out, ix->xyz, (size_t)ix->n * 24, cudaMemcpyDeviceToDevice, as_stream(stream)));
  return G2P_OK;
}

int g2p_index_max_nn_dist(g2p_index* ix, double* dmax, void* stream) {
  if (!ix) return set_error(G2P_ERR_INVALID, "g2p_index_max_nn_dist: null argument");
  return g2p_index_max_nn_dist_part(ix, 0, ix->n, dmax, stream);
}

int g2p_index_max_nn_dist_part(g2p_index* ix, int64_t first, int64_t last, double* dmax, void* stream) {
  if (!ix || !dmax) return set_error(G2P_ERR_INVALID, "g2p_index_max_nn_dist: null argument");
  if (first < 0 || last > ix->n || first > last) return set_error(G2P_ERR_INVALID, "g2p_index_max_nn_dist_part: bad range");
  if (first == last) {
    *dmax = 0.0;
    return G2P_OK;
  }
  cudaStream_t st = as_stream(stream);
  double* d_max = nullptr;
  G2P_CUDA(pool_alloc(&d_max, sizeof(double), st));
  G2P_CUDA(cudaMemsetAsync(d_max, 0, sizeof(double), st));
  KnnArgs a{};
  a.r0_scale = env_double("G2P_KNN_R0", 1.0);
  a.ix = view_of(ix);
  a.m = last - first;
  a.q0 = first;
  a.k = 2;
  a.mode = G2P_MODE_RAW;
  a.stop2 = __builtin_huge_val();
  a.dmax2_out = d_max;
  int rc = launch_knn(a, QUERY_SELF, st);
  double d2 = 0.0;
  if (rc == G2P_OK) {
    cudaError_t e = cudaMemcpyAsync(&d2, d_max, sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = cuda_fail(e, "max_nn_dist readback", __FILE__, __LINE__);
  }
  pool_free(d_max, sizeof(double));
  if (rc != G2P_OK) return rc;
  *dmax = sqrt(d2);  // sqrt is monotone and correctly rounded: sqrt(max d2) == max sqrt(d2)
  return G2P_OK;
}

int g2p_knn(const g2p_index* ix, const void* tlon, const void* tlat, const double* txyz, int64_t m, int dtype,
            double radius, const g2p_rotation* rot, int k, int mode, double mub, int32_t* idx_out, double* coef_out,
            uint8_t* flags_out, double* xyz_out, void* stream) {
  return g2p_knn_block(ix, tlon, tlat, txyz, m, dtype, radius, rot, k, mode, mub, idx_out, coef_out, m, 0, flags_out,
                       xyz_out, stream);
}

int g2p_knn_block(const g2p_index* ix, const void* tlon, const void* tlat, const double* txyz, int64_t m, int dtype,
                  double radius, const g2p_rotation* rot, int k, int mode, double mub, int32_t* idx_out, double* coef_out,
                  int64_t out_stride, int64_t out_offset, uint8_t* flags_out, double* xyz_out, void* stream) {
  if (out_stride < m || out_offset < 0 || out_offset + m > out_stride)
    return set_error(G2P_ERR_INVALID, "g2p_knn_block: the block [%lld, %lld) does not fit rows of %lld", (long long)out_offset,
                     (long long)(out_offset + m), (long long)out_stride);
  if (!ix || !idx_out || !coef_out) return set_error(G2P_ERR_INVALID, "g2p_knn: null index/outputs");
  if (!txyz && (!tlon || !tlat)) return set_error(G2P_ERR_INVALID, "g2p_knn: give tlon+tlat or txyz");
  if (k < 1 || k > 16) return set_error(G2P_ERR_INVALID, "g2p_knn: k=%d out of range 1..16", k);
  if (dtype != G2P_F64 && dtype != G2P_F32) return set_error(G2P_ERR_INVALID, "g2p_knn: bad dtype");
  if (mode != G2P_MODE_RAW && mode != G2P_MODE_NEAREST && mode != G2P_MODE_INVDIST)
    return set_error(G2P_ERR_INVALID, "g2p_knn: bad mode %d", mode);
  if (mode == G2P_MODE_NEAREST && k != 1) return set_error(G2P_ERR_INVALID, "g2p_knn: NEAREST needs k == 1");
  if (m < 0) return set_error(G2P_ERR_INVALID, "g2p_knn: m < 0");
  if (m == 0) return G2P_OK;
  KnnArgs a{};
  a.r0_scale = env_double("G2P_KNN_R0", 1.0);
  a.ix = view_of(ix);
  a.tlon = tlon;
  a.tlat = tlat;
  a.txyz = txyz;
  a.m = m;
  a.os = out_stride;
  a.oo = out_offset;
  a.rot = make_rot(rot);
  a.xyz_radius = radius > 0.0 ? radius : (ix->user_radius > 0.0 ? ix->user_radius : ix->radius);
  a.k = k;
  a.mode = mode;
  a.mub = mub;
  a.stop2 = (mode == G2P_MODE_INVDIST && mub > 0.0 && mub < 1e300) ? (mub * (1.0 + 1e-6)) * (mub * (1.0 + 1e-6))
                                                                    : __builtin_huge_val();
  a.idx_out = idx_out;
  a.coef_out = coef_out;
  a.flags_out = flags_out;
  a.xyz_out = xyz_out;
  int qsrc = txyz ? QUERY_XYZ : (dtype == G2P_F32 ? QUERY_LATLON_F32 : QUERY_LATLON_F64);
  // queries far from every source point (a regional source, a larger target) are listed by the main kernel and answered
  // by the coarse-to-fine far pass; G2P_KNN_FAR=0 keeps the exhaustive scan in place (experiments)
  static const bool far_pass = env_double("G2P_KNN_FAR", 1.0) != 0.0;
  uint32_t* d_far = nullptr;
  unsigned int* d_far_n = nullptr;
  const bool use_far = far_pass && ix->blocks && ix->n_blocks > 0 && m < 0xffffffffLL;
  if (use_far) {
    G2P_CUDA(pool_alloc(&d_far, (size_t)m * 4, as_stream(stream)));
    G2P_CUDA(pool_alloc(&d_far_n, sizeof(unsigned int), as_stream(stream)));
    G2P_CUDA(cudaMemsetAsync(d_far_n, 0, sizeof(unsigned int), as_stream(stream)));
    a.far_rows = d_far;
    a.far_count = d_far_n;
  }
  auto run = [&]() {
    int rc = launch_knn(a, qsrc, as_stream(stream));
    if (rc == G2P_OK && use_far) rc = launch_far(a, qsrc, as_stream(stream));
    if (use_far) {   // stream order protects the reuse of pooled buffers
      pool_free(d_far, (size_t)m * 4);
      pool_free(d_far_n, sizeof(unsigned int));
    }
    return rc;
  };
  if (getenv("G2P_KNN_STATS")) {  // diagnostics: search passes and candidates per query, on stderr (synchronises)
    unsigned long long* d_stats = nullptr;
    G2P_CUDA(cudaMalloc(&d_stats, 16));
    G2P_CUDA(cudaMemsetAsync(d_stats, 0, 16, as_stream(stream)));
    a.stats = d_stats;
    const int rc = run();
    unsigned long long h[2] = {0, 0};
    cudaMemcpyAsync(h, d_stats, 16, cudaMemcpyDeviceToHost, as_stream(stream));
    cudaStreamSynchronize(as_stream(stream));
    cudaFree(d_stats);
    fprintf(stderr, "g2p_knn: %lld queries k=%d: %.3f extra passes and %.1f candidates per query\n", (long long)m, k,
            (double)h[0] / (double)m, (double)h[1] / (double)m);
    return rc;
  }
  return run();
}

int g2p_weights(int mode, int k, int64_t m, int64_t n_src, double mub, int dtype, int32_t* idx, double* coef,
                void* stream) {
  if (!idx || !coef) return set_error(G2P_ERR_INVALID, "g2p_weights: null pointer");
  if (mode != G2P_MODE_NEAREST && mode != G2P_MODE_INVDIST) return set_error(G2P_ERR_INVALID, "g2p_weights: bad mode");
  if (k < 1 || k > 16 || (mode == G2P_MODE_NEAREST && k != 1)) return set_error(G2P_ERR_INVALID, "g2p_weights: bad k");
  if (m <= 0) return G2P_OK;
  cudaStream_t st = as_stream(stream);
  if (dtype == G2P_F32)
    weights_kernel<true><<<grid_for(m, 256), 256, 0, st>>>(mode, k, m, (int32_t)n_src, mub, idx, coef);
  else
    weights_kernel<false><<<grid_for(m, 256), 256, 0, st>>>(mode, k, m, (int32_t)n_src, mub, idx, coef);
  G2P_LAUNCHED();
  return G2P_OK;
}

int g2p_table_compact(const int32_t* idx, int64_t m, int k, int64_t n, int32_t* touched_out, int32_t* cidx_out,
                      int64_t* nc_out, void* stream) {
  if (!idx || !touched_out || !cidx_out || !nc_out) return set_error(G2P_ERR_INVALID, "g2p_table_compact: null pointer");
  if (m < 0 || k < 1 || k > 16 || n < 1 || n > 0x7ffffff0LL) return set_error(G2P_ERR_INVALID, "g2p_table_compact: bad sizes");
  cudaStream_t st = as_stream(stream);
  const int64_t total = m * k;
  const int nb = (int)ceil_div(n, kScanBlock);
  uint32_t *d_used = nullptr, *d_pos = nullptr, *d_bsums = nullptr;
  G2P_CUDA(pool_alloc(&d_used, (size_t)n * 4, st));
  G2P_CUDA(pool_alloc(&d_pos, (size_t)(n + 1) * 4, st));
  G2P_CUDA(pool_alloc(&d_bsums, (size_t)nb * 4, st));
  G2P_CUDA(cudaMemsetAsync(d_used, 0, (size_t)n * 4, st));
  if (total > 0) {
    mark_touched_kernel<<<grid_for(total, 256), 256, 0, st>>>(idx, total, n, d_used);
    G2P_LAUNCHED();
  }
  scan_blocks_kernel<<<nb, kScanThreads, 0, st>>>(d_used, d_pos, n, d_bsums);
  G2P_LAUNCHED();
  scan_sums_kernel<<<1, 1024, 0, st>>>(d_bsums, nb);
  G2P_LAUNCHED();
  scan_add_kernel<<<nb, kScanThreads, 0, st>>>(d_pos, n, d_bsums, nullptr, 0u);
  G2P_LAUNCHED();
  uint32_t last_pos = 0, last_used = 0;
  G2P_CUDA(cudaMemcpyAsync(&last_pos, d_pos + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  G2P_CUDA(cudaMemcpyAsync(&last_used, d_used + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  G2P_CUDA(cudaStreamSynchronize(st));
  const int64_t nc = (int64_t)last_pos + last_used;
  list_touched_kernel<<<grid_for(n, 256), 256, 0, st>>>(d_used, d_pos, n, touched_out);
  G2P_LAUNCHED();
  if (total > 0) {
    renumber_kernel<<<grid_for(total, 256), 256, 0, st>>>(idx, total, n, d_pos, (int32_t)nc, cidx_out);
    G2P_LAUNCHED();
  }
  G2P_CUDA(cudaStreamSynchronize(st));
  pool_free(d_used, (size_t)n * 4);
  pool_free(d_pos, (size_t)(n + 1) * 4);
  pool_free(d_bsums, (size_t)nb * 4);
  *nc_out = nc;
  return G2P_OK;
}

}  // extern "C"
