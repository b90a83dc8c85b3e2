// libm_exact.cuh -- device replicas of the three glibc entry points that the
// reference's fmPLL reaches (src/filter.cpp:58 atan2f, :69-70 sincosf, :71 cosf
// once g++ -O3 has lowered std::atan2/std::cos/std::sin on floats).
//
// The PLL is numerically ill-conditioned (hard part 1): its phase
// argument is rounded to float while it grows without bound, so a 1-ulp
// difference in any transcendental changes later outputs by far more than the
// parity tolerance.  These functions therefore perform, operation by operation,
// what glibc 2.39 on x86-64 performs:
//   atan2f / atanf  : fdlibm single-precision code (e_atan2f.c, s_atanf.c); the
//                     library build has no FMA variant of these, so every
//                     multiply and add rounds separately;
//   sincosf / cosf  : the double-precision polynomial code (s_sincosf.c,
//                     sincosf_poly.h) in the FMA ifunc variant selected on every
//                     AVX2+FMA host; fma placement as in that variant.
// The CPU restatement with the same structure is oracle/fm_oracle.c, which is
// pinned against the host libm.
#pragma once

#include "common.cuh"

namespace sdr {

__device__ __forceinline__ float atanf_glibc(float x) {
  const float hi0 = 4.6364760399e-01f, hi1 = 7.8539812565e-01f, hi2 = 9.8279368877e-01f,
              hi3 = 1.5707962513e+00f;
  const float lo0 = 5.0121582440e-09f, lo1 = 3.7748947079e-08f, lo2 = 3.4473217170e-08f,
              lo3 = 7.5497894159e-08f;
  const float a0 = 3.3333334327e-01f, a1 = -2.0000000298e-01f, a2 = 1.4285714924e-01f,
              a3 = -1.1111110449e-01f, a4 = 9.0908870101e-02f, a5 = -7.6918758452e-02f,
              a6 = 6.6610731184e-02f, a7 = -5.8335702866e-02f, a8 = 4.9768779427e-02f,
              a9 = -3.6531571299e-02f, a10 = 1.6285819933e-02f;
  int32_t hx = __float_as_int(x);
  int32_t ix = hx & 0x7fffffff;
  float hi = 0.0f, lo = 0.0f;
  bool reduced = true;
  if (ix >= 0x4c000000) {  // |x| >= 2^25
    if (ix > 0x7f800000) return xadd(x, x);
    return (hx > 0) ? xadd(hi3, lo3) : xsub(-hi3, lo3);
  }
  if (ix < 0x3ee00000) {    // |x| < 0.4375
    if (ix < 0x31000000) {  // |x| < 2^-29
      if (xadd(1.0e30f, x) > 1.0f) return x;
    }
    reduced = false;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {    // |x| < 1.1875
      if (ix < 0x3f300000) {  // 7/16 <= |x| < 11/16
        hi = hi0; lo = lo0;
        x = xdiv(xsub(xmul(2.0f, x), 1.0f), xadd(2.0f, x));
      } else {                // 11/16 <= |x| < 19/16
        hi = hi1; lo = lo1;
        x = xdiv(xsub(x, 1.0f), xadd(x, 1.0f));
      }
    } else {
      if (ix < 0x401c0000) {  // |x| < 2.4375
        hi = hi2; lo = lo2;
        x = xdiv(xsub(x, 1.5f), xadd(1.0f, xmul(1.5f, x)));
      } else {
        hi = hi3; lo = lo3;
        x = xdiv(-1.0f, x);
      }
    }
  }
  float z = xmul(x, x);
  float w = xmul(z, z);
  float s1 = xmul(z, xadd(a0, xmul(w, xadd(a2, xmul(w, xadd(a4, xmul(w, xadd(a6, xmul(w, xadd(a8, xmul(w, a10)))))))))));
  float s2 = xmul(w, xadd(a1, xmul(w, xadd(a3, xmul(w, xadd(a5, xmul(w, xadd(a7, xmul(w, a9)))))))));
  float t = xmul(x, xadd(s1, s2));
  if (!reduced) return xsub(x, t);
  z = xsub(hi, xsub(xsub(t, lo), x));
  return (hx < 0) ? -z : z;
}

__device__ __forceinline__ float atan2f_glibc(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f,
              pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
  int32_t hx = __float_as_int(x), hy = __float_as_int(y);
  int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return xadd(x, y);
  if (hx == 0x3f800000) return atanf_glibc(y);
  int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    if (m < 2) return y;
    return (m == 2) ? xadd(pi, tiny) : xsub(-pi, tiny);
  }
  if (ix == 0) return (hy < 0) ? xsub(-pi_o_2, tiny) : xadd(pi_o_2, tiny);
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      switch (m) {
        case 0: return xadd(pi_o_4, tiny);
        case 1: return xsub(-pi_o_4, tiny);
        case 2: return xadd(xmul(3.0f, pi_o_4), tiny);
        default: return xsub(xmul(-3.0f, pi_o_4), tiny);
      }
    }
    switch (m) {
      case 0: return 0.0f;
      case 1: return -0.0f;
      case 2: return xadd(pi, tiny);
      default: return xsub(-pi, tiny);
    }
  }
  if (iy == 0x7f800000) return (hy < 0) ? xsub(-pi_o_2, tiny) : xadd(pi_o_2, tiny);
  int32_t k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = xadd(pi_o_2, xmul(0.5f, pi_lo));
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = atanf_glibc(fabsf(xdiv(y, x)));
  switch (m) {
    case 0: return z;
    case 1: return __int_as_float(__float_as_int(z) ^ (int32_t)0x80000000);
    case 2: return xsub(pi, xsub(z, pi_lo));
    default: return xsub(xsub(z, pi_lo), pi);
  }
}

// 4/pi as overlapping 32-bit windows (glibc __inv_pio4).
static __device__ __constant__ uint32_t k_inv_pio4[24] = {
    0xa2,       0xa2f9,     0xa2f983,   0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529,
    0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0,
    0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041};

// Argument reduction shared by sincosf/cosf (s_sincosf.h reduce_fast/reduce_large).
// Returns false for the trivial cases (tiny, inf/nan) which the callers finish.
// n: quadrant count whose low bit swaps sin/cos; q: index into sign[]/table.
struct TrigRed {
  double x;
  int n, q;
};

__device__ __forceinline__ int trig_reduce(float y, TrigRed &r) {
  const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
  uint32_t bits = __float_as_uint(y);
  uint32_t top = (bits >> 20) & 0x7ff;
  r.x = (double)y;
  r.n = 0;
  r.q = 0;
  if (top < 0x3f4) return (top < 0x398) ? 1 : 0;  // 1: |y| < 2^-12
  if (top < 0x42f) {                              // |y| < 120
    double t = __dmul_rn(r.x, hpi_inv);
    r.n = (__double2int_rz(t) + 0x800000) >> 24;
    r.x = __fma_rn(-(double)r.n, hpi, r.x);
    r.q = r.n;
    return 0;
  }
  if (top < 0x7f8) {
    const uint32_t *arr = &k_inv_pio4[(bits >> 26) & 15];
    int shift = (bits >> 23) & 7;
    uint32_t xi = ((bits & 0xffffffu) | 0x800000u) << shift;
    uint64_t res0 = (uint32_t)(xi * arr[0]);
    uint64_t res1 = (uint64_t)xi * arr[4];
    uint64_t res2 = (uint64_t)xi * arr[8];
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    uint64_t nn = (res0 + (1ULL << 61)) >> 62;
    res0 -= nn << 62;
    r.x = __dmul_rn((double)(int64_t)res0, 0x1.921FB54442D18p-62);
    r.n = (int)nn;
    r.q = r.n + (int)(bits >> 31);
    return 0;
  }
  return 2;  // inf / nan
}

// sincosf_poly.h sine polynomial; sign[q&3] = {1,-1,-1,1} multiplies x first.
__device__ __forceinline__ float trig_sin_poly(double x, double x2, int q) {
  const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
  double xs = ((q + 1) & 2) ? -x : x;
  double x3 = __dmul_rn(x2, xs);
  double s1 = __fma_rn(x2, S3, S2);
  double x5 = __dmul_rn(x2, x3);
  double s = __fma_rn(x3, S1, xs);
  return __double2float_rn(__fma_rn(s1, x5, s));
}
// cosine polynomial; table 1 (q&2) holds the negated coefficients.
__device__ __forceinline__ float trig_cos_poly(double x2, int q) {
  const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5,
               C3 = -0x1.6c087e89a359dp-10, C4 = 0x1.99343027bf8c3p-16;
  const bool ng = (q & 2) != 0;
  double x4 = __dmul_rn(x2, x2);
  double c1 = __fma_rn(x2, ng ? -C1 : C1, ng ? -C0 : C0);
  double c2 = __fma_rn(x2, ng ? -C4 : C4, ng ? -C3 : C3);
  double x6 = __dmul_rn(x2, x4);
  double c = __fma_rn(x4, ng ? -C2 : C2, c1);
  return __double2float_rn(__fma_rn(c2, x6, c));
}

// sin and cos of a float, bit-identical to glibc's sincosf.
__device__ __forceinline__ void sincosf_glibc(float y, float &sinv, float &cosv) {
  TrigRed r;
  int special = trig_reduce(y, r);
  if (special == 1) {
    sinv = y;
    cosv = 1.0f;
    return;
  }
  if (special == 2) {
    sinv = cosv = xsub(y, y);
    return;
  }
  double x2 = __dmul_rn(r.x, r.x);
  float fs = trig_sin_poly(r.x, x2, r.q);
  float fc = trig_cos_poly(x2, r.q);
  sinv = (r.n & 1) ? fc : fs;
  cosv = (r.n & 1) ? fs : fc;
}

// glibc cosf: the same reduction; only the polynomial that lands in "cos" runs.
__device__ __forceinline__ float cosf_glibc(float y) {
  TrigRed r;
  int special = trig_reduce(y, r);
  if (special == 1) return 1.0f;
  if (special == 2) return xsub(y, y);
  double x2 = __dmul_rn(r.x, r.x);
  return (r.n & 1) ? trig_sin_poly(r.x, x2, r.q) : trig_cos_poly(x2, r.q);
}


// ---------------------------------------------------------------------------
// Branch-free forms for the PLL's inner loop.  One lane walks one capture
// sequentially, so the loop is bound by the LATENCY of its dependent chain, not by
// issue slots; removing the data-dependent branches lets the scheduler overlap the
// independent pieces (the two Horner chains of atanf, the sine and cosine
// polynomials, the NCO output of the previous sample).  Every value is produced by
// the same operations as in the branchy forms above; only the selection of
// constants and results is done with selects instead of jumps.
// ---------------------------------------------------------------------------

// atan2f for the common case: x, y finite and non-zero, x != 1.0f, exponents within 2^60
// of each other, |y/x| < 2^25.  Returns false otherwise (`out` is then meaningless); the
// caller then falls back to atan2f_glibc.  The validity test is evaluated LAST: the common path
// runs on whatever arrives (a zero, infinity or NaN only produces a value that is discarded), so
// the branch on it -- which needs the first quotient -- does not sit in the middle of the PLL's
// dependency chain (305 -> 280 cycles, tools/ubench_pll_chain.cu).
__device__ __forceinline__ bool atan2f_common(float y, float x, float &out) {
  const float pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
  const int32_t hx = __float_as_int(x), hy = __float_as_int(y);
  const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  const int32_t k = (iy - ix) >> 23;
  const float t = fabsf(xdiv(y, x));
  const int32_t it = __float_as_int(t);
  const bool common = (iy != 0) & (ix != 0) & (ix < 0x7f800000) & (iy < 0x7f800000) &
                      (hx != 0x3f800000) & (k <= 60) & (k >= -60) & (it < 0x4c000000);
  // atanf(t), t >= 0: argument reduction x = num/den with per-range constants
  //   range        num                den               hi/lo
  //   t < 7/16     t                  1                 0
  //   < 11/16      2t - 1             2 + t             atan(0.5)
  //   < 19/16      t - 1              t + 1             atan(1)
  //   < 39/16      t - 1.5            1 + 1.5t          atan(1.5)
  //   else         -1                 t                 atan(inf)
  const bool r0 = it < 0x3ee00000, r1 = it < 0x3f300000, r2 = it < 0x3f980000, r3 = it < 0x401c0000;
  const float na = r0 ? 1.0f : r1 ? 2.0f : r3 ? 1.0f : 0.0f;
  const float nb = r0 ? 0.0f : r2 ? -1.0f : r3 ? -1.5f : -1.0f;
  const float da = r0 ? 0.0f : r2 ? 1.0f : r3 ? 1.5f : 1.0f;
  const float db = r0 ? 1.0f : r1 ? 2.0f : r3 ? 1.0f : 0.0f;
  const float hi = r0 ? 0.0f : r1 ? 4.6364760399e-01f : r2 ? 7.8539812565e-01f : r3 ? 9.8279368877e-01f : 1.5707962513e+00f;
  const float lo = r0 ? 0.0f : r1 ? 5.0121582440e-09f : r2 ? 3.7748947079e-08f : r3 ? 3.4473217170e-08f : 7.5497894159e-08f;
  const float num = xadd(xmul(na, t), nb);
  const float den = xadd(xmul(da, t), db);
  const float xr = xdiv(num, den);
  const float a0 = 3.3333334327e-01f, a1 = -2.0000000298e-01f, a2 = 1.4285714924e-01f,
              a3 = -1.1111110449e-01f, a4 = 9.0908870101e-02f, a5 = -7.6918758452e-02f,
              a6 = 6.6610731184e-02f, a7 = -5.8335702866e-02f, a8 = 4.9768779427e-02f,
              a9 = -3.6531571299e-02f, a10 = 1.6285819933e-02f;
  const float z = xmul(xr, xr);
  const float w = xmul(z, z);
  const float s1 = xmul(z, xadd(a0, xmul(w, xadd(a2, xmul(w, xadd(a4, xmul(w, xadd(a6, xmul(w, xadd(a8, xmul(w, a10)))))))))));
  const float s2 = xmul(w, xadd(a1, xmul(w, xadd(a3, xmul(w, xadd(a5, xmul(w, xadd(a7, xmul(w, a9)))))))));
  const float ts = xmul(xr, xadd(s1, s2));
  // r0: x - x*(s1+s2); otherwise hi - ((x*(s1+s2) - lo) - x).  With hi = lo = 0 the second
  // form evaluates to the first one bit for bit (negation commutes with rounding).
  const float zz = xsub(hi, xsub(xsub(ts, lo), xr));
  // quadrant (e_atan2f.c switch on m = 2*sign(x) + sign(y))
  const float zl = xsub(zz, pi_lo);
  const bool xneg = hx < 0, yneg = hy < 0;
  const float pos = xneg ? xsub(pi, zl) : zz;
  const float neg = xneg ? xsub(zl, pi) : __int_as_float(__float_as_int(zz) ^ (int32_t)0x80000000);
  out = yneg ? neg : pos;
  return common;
}

// Branch-free argument reduction: all three glibc paths are evaluated and the one
// that applies is selected.  `special`: 0 normal, 1 |y| < 2^-12, 2 inf/nan.
__device__ __forceinline__ int trig_reduce_bf(float y, TrigRed &r) {
  const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
  const uint32_t bits = __float_as_uint(y);
  const uint32_t top = (bits >> 20) & 0x7ff;
  const double xd = (double)y;
  // reduce_fast (|y| < 120)
  const double tf = __dmul_rn(xd, hpi_inv);
  const int nf = (__double2int_rz(tf) + 0x800000) >> 24;
  const double xf = __fma_rn(-(double)nf, hpi, xd);
  // reduce_large
  const uint32_t *arr = &k_inv_pio4[(bits >> 26) & 15];
  const int shift = (bits >> 23) & 7;
  const uint32_t xi = ((bits & 0xffffffu) | 0x800000u) << shift;
  uint64_t res0 = (uint32_t)(xi * arr[0]);
  const uint64_t res1 = (uint64_t)xi * arr[4];
  const uint64_t res2 = (uint64_t)xi * arr[8];
  res0 = (res2 >> 32) | (res0 << 32);
  res0 += res1;
  const uint64_t nn = (res0 + (1ULL << 61)) >> 62;
  res0 -= nn << 62;
  const double xl = __dmul_rn((double)(int64_t)res0, 0x1.921FB54442D18p-62);
  const bool small = top < 0x3f4, fast = top < 0x42f;
  r.x = small ? xd : fast ? xf : xl;
  r.n = small ? 0 : fast ? nf : (int)nn;
  r.q = small ? 0 : fast ? nf : (int)nn + (int)(bits >> 31);
  return (top < 0x398) ? 1 : (top >= 0x7f8) ? 2 : 0;
}

__device__ __forceinline__ void sincosf_glibc_bf(float y, float &sinv, float &cosv) {
  TrigRed r;
  const int special = trig_reduce_bf(y, r);
  const double x2 = __dmul_rn(r.x, r.x);
  const float fs = trig_sin_poly(r.x, x2, r.q);
  const float fc = trig_cos_poly(x2, r.q);
  float s = (r.n & 1) ? fc : fs;
  float c = (r.n & 1) ? fs : fc;
  if (special == 1) { s = y; c = 1.0f; }
  if (special == 2) { s = c = xsub(y, y); }
  sinv = s;
  cosv = c;
}

__device__ __forceinline__ float cosf_glibc_bf(float y) {
  float s, c;
  sincosf_glibc_bf(y, s, c);
  return c;
}

}  // namespace sdr
