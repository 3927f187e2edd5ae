// float32 sin / cos with the bits of glibc's sinf / cosf (glibc >= 2.28, sysdeps/ieee754/flt-32/s_sinf.c,
// s_cosf.c, sincosf.h, sincosf_data.c - the double-precision-polynomial algorithm of ARM's optimized routines).
//
// Why: the reference evaluates `to_3d` on PCRaster REAL4 maps through numexpr's float32 opcodes, which call the
// C library's sinf / cosf (scipy_interpolation_lib.py:667-694).  glibc's functions are NOT correctly rounded
// (0.56 ulp worst case), so rounding a float64 sine reproduces them in 98.7 % of the values only; the targets of the
// other 1.3 % moved by ~1 m.  This header restates glibc's published algorithm operation for operation - reduction by
// one scaled multiply (|x| < 120) or by the 192-bit 4/pi table (larger), one degree-7 / degree-8 double polynomial -
// so that the float32 results are bit-identical.  It compiles as CUDA device code and as plain host C++ (the host
// build is what tools/sincosf_check.cpp runs against the C library over all 2^32 arguments).
//
// glibc builds s_sinf.c twice on x86_64 (sysdeps/x86_64/fpu/multiarch: __sinf_fma with -mfma -mavx2, __sinf_sse2)
// and picks one at load time; with -ffp-contract=fast every `a + b * c` of the source becomes one fused
// multiply-add in the first variant.  `FMA` selects which of the two is restated.  Checked over all 2^32 arguments
// against glibc 2.39 on an FMA host (tools/sincosf_check.cpp): the FMA restatement differs in 0 results of sinf and
// 0 of cosf; the non-FMA restatement differs from the FMA build in 12 (sinf) + 22 (cosf) of the 2^32 arguments.
#ifndef G2P_SINCOSF_H
#define G2P_SINCOSF_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define G2P_SCF_FN __device__ __forceinline__
#define G2P_SCF_MUL(a, b) __dmul_rn((a), (b))
#define G2P_SCF_ADD(a, b) __dadd_rn((a), (b))
#define G2P_SCF_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#include <math.h>
#if defined(__CUDACC__)
#define G2P_SCF_FN __host__ __forceinline__
#else
#define G2P_SCF_FN static inline
#endif
// host build: compile with -ffp-contract=off so that only the explicit fma() calls fuse
#define G2P_SCF_MUL(a, b) ((a) * (b))
#define G2P_SCF_ADD(a, b) ((a) + (b))
#define G2P_SCF_FMA(a, b, c) fma((a), (b), (c))
#endif

namespace g2p_scf {

// a + b*c as the selected glibc build evaluates it
template <bool FMA>
G2P_SCF_FN double madd(double a, double b, double c) {
  return FMA ? G2P_SCF_FMA(b, c, a) : G2P_SCF_ADD(a, G2P_SCF_MUL(b, c));
}

G2P_SCF_FN uint32_t as_u32(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(x);
#else
  uint32_t u;
  memcpy(&u, &x, 4);
  return u;
#endif
}
G2P_SCF_FN uint32_t abstop12(float x) { return (as_u32(x) >> 20) & 0x7ffu; }

// sincosf_data.c: the two coefficient sets differ in the sign of the cosine polynomial only (`negcos`)
template <bool FMA>
G2P_SCF_FN float sinf_poly(double x, double x2, bool negcos, int n) {
  const double sg = negcos ? -1.0 : 1.0;
  if ((n & 1) == 0) {
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    const double x3 = G2P_SCF_MUL(x, x2);
    const double s1 = madd<FMA>(S2, x2, S3);
    const double x7 = G2P_SCF_MUL(x3, x2);
    const double s = madd<FMA>(x, x3, S1);
    return (float)madd<FMA>(s, x7, s1);
  } else {
    const double C0 = sg * 0x1p0, C1 = sg * -0x1.ffffffd0c621cp-2, C2 = sg * 0x1.55553e1068f19p-5,
                 C3 = sg * -0x1.6c087e89a359dp-10, C4 = sg * 0x1.99343027bf8c3p-16;
    const double x4 = G2P_SCF_MUL(x2, x2);
    const double c2 = madd<FMA>(C3, x2, C4);
    const double c1 = madd<FMA>(C0, x2, C1);
    const double x6 = G2P_SCF_MUL(x4, x2);
    const double c = madd<FMA>(c1, x4, C2);
    return (float)madd<FMA>(c, x6, c2);
  }
}

// reduce_fast (the !TOINT_INTRINSICS branch, which is what x86_64 builds): hpi_inv is 2/pi * 2^24
template <bool FMA>
G2P_SCF_FN double reduce_fast(double x, int* np) {
  const double r = G2P_SCF_MUL(x, 0x1.45F306DC9C883p+23);
  const int n = ((int32_t)r + 0x800000) >> 24;
  *np = n;
  // x - n * hpi
  return FMA ? G2P_SCF_FMA(-(double)n, 0x1.921FB54442D18p0, x) : G2P_SCF_ADD(x, -G2P_SCF_MUL((double)n, 0x1.921FB54442D18p0));
}

// 4/pi as 32-bit windows at byte steps (__inv_pio4): word i holds bits [8i-24 .. 8i+8) of the binary expansion
G2P_SCF_FN uint32_t inv_pio4(int i) {
  // the byte string 00 00 00 a2 f9 83 6e 4e 44 15 29 fc 27 57 d1 f5 34 dd c0 db 62 95 99 3c 43 90 41 (4/pi = 0x1.45f306dc9c88...)
  // held in registers (no table in local memory): word i = bytes i .. i+3
  const uint64_t A0 = 0x000000a2f9836e4eull, A1 = 0x441529fc2757d1f5ull, A2 = 0x34ddc0db6295993cull,
                 A3 = 0x4390410000000000ull;
  const int q = i >> 3, r = (i & 7) * 8;
  const uint64_t hi = q == 0 ? A0 : q == 1 ? A1 : A2;
  const uint64_t lo = q == 0 ? A1 : q == 1 ? A2 : A3;
  const uint64_t w = r ? ((hi << r) | (lo >> (64 - r))) : hi;
  return (uint32_t)(w >> 32);
}

// reduce_large: exact 2.62 fixed-point x mod pi/2 from a 32 x 96 -> 128 bit multiply
G2P_SCF_FN double reduce_large(uint32_t xi, int* np) {
  const int a = (int)((xi >> 26) & 15u);
  const int shift = (int)((xi >> 23) & 7u);
  xi = (xi & 0xffffffu) | 0x800000u;
  xi <<= shift;
  uint64_t res0 = (uint64_t)(uint32_t)(xi * inv_pio4(a));  // 32-bit product, as in glibc (`xi * arr[0]` is uint32_t)
  const uint64_t res1 = (uint64_t)xi * inv_pio4(a + 4);
  const uint64_t res2 = (uint64_t)xi * inv_pio4(a + 8);
  res0 = (res2 >> 32) | (res0 << 32);
  res0 += res1;
  const uint64_t n = (res0 + (1ULL << 61)) >> 62;
  res0 -= n << 62;
  const double x = (double)(int64_t)res0;
  *np = (int)n;
  return G2P_SCF_MUL(x, 0x1.921FB54442D18p-62);  // pi63 = 2 pi * 2^-64
}

// COS = false: sinf(y); COS = true: cosf(y)
template <bool FMA, bool COS>
G2P_SCF_FN float sincosf_glibc(float y) {
  double x = (double)y;
  int n;
  const uint32_t top = abstop12(y);
  if (top < 0x3f4u) {  // abstop12(pio4 = 0x1.921FB6p-1f)
    const double x2 = G2P_SCF_MUL(x, x);
    if (top < 0x398u)  // abstop12(0x1p-12f)
      return COS ? 1.0f : y;
    return sinf_poly<FMA>(x, x2, false, COS ? 1 : 0);
  }
  if (top < 0x42fu) {  // abstop12(120.0f)
    x = reduce_fast<FMA>(x, &n);
    const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;  // sign[] = {1, -1, -1, 1}
    return sinf_poly<FMA>(G2P_SCF_MUL(x, s), G2P_SCF_MUL(x, x), (n & 2) != 0, COS ? (n ^ 1) : n);
  }
  if (top < 0x7f8u) {  // finite
    const uint32_t xi = as_u32(y);
    const int sign = (int)(xi >> 31);
    x = reduce_large(xi, &n);
    const int q = (n + sign) & 3;
    const double s = (q == 1 || q == 2) ? -1.0 : 1.0;
    return sinf_poly<FMA>(G2P_SCF_MUL(x, s), G2P_SCF_MUL(x, x), ((n + sign) & 2) != 0, COS ? (n ^ 1) : n);
  }
  return y - y;  // __math_invalidf: NaN for +-inf and NaN
}

}  // namespace g2p_scf

#endif  // G2P_SINCOSF_H
