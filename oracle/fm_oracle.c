/*
 * fm_oracle.c -- plain-C CPU restatement of the reference FM receiver DSP chain.
 *
 * TEST INFRASTRUCTURE ONLY (see fm_oracle.h).  Written from the behaviour of the
 * reference, not copied from it; every function cites the reference lines it
 * restates.  Compile WITHOUT -ffast-math and WITH -ffp-contract=off: the
 * arithmetic order and the separate rounding of every multiply and add are part
 * of the contract (the reference is built for baseline x86-64, i.e. scalar
 * mulss/addss, no FMA).
 *
 * Parity: PINNED against oracle/_ref/libfmref.so (the reference's own
 * filter.cpp + iofunc.cpp compiled unmodified) and tests/golden/.
 */
#include "fm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI 3.14159265358979323846 /* include/dy4.h:23 */

/* ======================================================================== */
/* libm replicas.                                                           */
/*                                                                          */
/* fmPLL (filter.cpp:58,69-71) calls std::atan2/cos/sin on floats; g++ -O3  */
/* emits calls to atan2f, sincosf and cosf.  libm is a third-party          */
/* dependency that is not vendored by the reference; on the build host it   */
/* resolves to glibc 2.39 (Ubuntu 2.39-0ubuntu8.5).  The algorithms below   */
/* restate glibc's published implementations:                               */
/*   atan2f/atanf : sysdeps/ieee754/flt-32/{e_atan2f.c,s_atanf.c} (fdlibm,  */
/*                  single precision, one rounding per operation)           */
/*   sincosf/cosf : sysdeps/ieee754/flt-32/{s_sincosf.c,s_cosf.c,sincosf_   */
/*                  poly.h} (ARM optimized routines; double precision       */
/*                  polynomial), in the x86_64 FMA ifunc variant that every */
/*                  AVX2+FMA host selects; the fma() placement below is the */
/*                  one found in that variant's machine code.               */
/* tests/test_oracle.py pins them against the host libm.                    */
/* ======================================================================== */

static inline uint32_t f2u(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static inline float u2f(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}

static const float k_atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f,
                                  1.5707962513e+00f};
static const float k_atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f,
                                  7.5497894159e-08f};
static const float k_aT[11] = {3.3333334327e-01f,  -2.0000000298e-01f, 1.4285714924e-01f,
                               -1.1111110449e-01f, 9.0908870101e-02f,  -7.6918758452e-02f,
                               6.6610731184e-02f,  -5.8335702866e-02f, 4.9768779427e-02f,
                               -3.6531571299e-02f, 1.6285819933e-02f};

static float orc_atanf(float x) {
  float w, s1, s2, z;
  int32_t hx = (int32_t)f2u(x);
  int32_t ix = hx & 0x7fffffff;
  int id;
  if (ix >= 0x4c000000) { /* |x| >= 2^25 */
    if (ix > 0x7f800000) return x + x;
    if (hx > 0) return k_atanhi[3] + k_atanlo[3];
    return -k_atanhi[3] - k_atanlo[3];
  }
  if (ix < 0x3ee00000) {   /* |x| < 0.4375 */
    if (ix < 0x31000000) { /* |x| < 2^-29 */
      if (1.0e30f + x > 1.0f) return x;
    }
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {   /* |x| < 1.1875 */
      if (ix < 0x3f300000) { /* 7/16 <= |x| < 11/16 */
        id = 0;
        x = (2.0f * x - 1.0f) / (2.0f + x);
      } else { /* 11/16 <= |x| < 19/16 */
        id = 1;
        x = (x - 1.0f) / (x + 1.0f);
      }
    } else {
      if (ix < 0x401c0000) { /* |x| < 2.4375 */
        id = 2;
        x = (x - 1.5f) / (1.0f + 1.5f * x);
      } else {
        id = 3;
        x = -1.0f / x;
      }
    }
  }
  z = x * x;
  w = z * z;
  s1 = z * (k_aT[0] + w * (k_aT[2] + w * (k_aT[4] + w * (k_aT[6] + w * (k_aT[8] + w * k_aT[10])))));
  s2 = w * (k_aT[1] + w * (k_aT[3] + w * (k_aT[5] + w * (k_aT[7] + w * k_aT[9]))));
  if (id < 0) return x - x * (s1 + s2);
  z = k_atanhi[id] - ((x * (s1 + s2) - k_atanlo[id]) - x);
  return (hx < 0) ? -z : z;
}

float orc_atan2f(float y, float x) {
  static const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f,
                     pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
  float z;
  int32_t hx = (int32_t)f2u(x), hy = (int32_t)f2u(y);
  int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  int32_t k, m;
  if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
  if (hx == 0x3f800000) return orc_atanf(y);
  m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    switch (m) {
      case 0:
      case 1:
        return y;
      case 2:
        return pi + tiny;
      default:
        return -pi - tiny;
    }
  }
  if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      switch (m) {
        case 0:
          return pi_o_4 + tiny;
        case 1:
          return -pi_o_4 - tiny;
        case 2:
          return 3.0f * pi_o_4 + tiny;
        default:
          return -3.0f * pi_o_4 - tiny;
      }
    } else {
      switch (m) {
        case 0:
          return 0.0f;
        case 1:
          return -0.0f;
        case 2:
          return pi + tiny;
        default:
          return -pi - tiny;
      }
    }
  }
  if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  k = (iy - ix) >> 23;
  if (k > 60)
    z = pi_o_2 + 0.5f * pi_lo;
  else if (hx < 0 && k < -60)
    z = 0.0f;
  else
    z = orc_atanf(fabsf(y / x));
  switch (m) {
    case 0:
      return z;
    case 1:
      return u2f(f2u(z) ^ 0x80000000u);
    case 2:
      return pi - (z - pi_lo);
    default:
      return (z - pi_lo) - pi;
  }
}

typedef struct {
  double sign[4];
  double hpi_inv, hpi;
  double c0, c1, c2, c3, c4;
  double s1, s2, s3;
} sc_tab;

static const sc_tab k_sc[2] = {
    {{1.0, -1.0, -1.0, 1.0},
     0x1.45F306DC9C883p+23,
     0x1.921FB54442D18p0,
     0x1p0,
     -0x1.ffffffd0c621cp-2,
     0x1.55553e1068f19p-5,
     -0x1.6c087e89a359dp-10,
     0x1.99343027bf8c3p-16,
     -0x1.555545995a603p-3,
     0x1.1107605230bc4p-7,
     -0x1.994eb3774cf24p-13},
    {{1.0, -1.0, -1.0, 1.0},
     0x1.45F306DC9C883p+23,
     0x1.921FB54442D18p0,
     -0x1p0,
     0x1.ffffffd0c621cp-2,
     -0x1.55553e1068f19p-5,
     0x1.6c087e89a359dp-10,
     -0x1.99343027bf8c3p-16,
     -0x1.555545995a603p-3,
     0x1.1107605230bc4p-7,
     -0x1.994eb3774cf24p-13}};

/* 4/pi as a 768-bit fixed-point table, one 32-bit window per byte offset. */
static const uint32_t k_inv_pio4[24] = {
    0xa2,       0xa2f9,     0xa2f983,   0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529,
    0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0,
    0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041};

static inline double sc_sin_poly(double xs, double x2, const sc_tab *p) {
  double x3 = x2 * xs;
  double s1 = fma(x2, p->s3, p->s2);
  double x5 = x2 * x3;
  double s = fma(x3, p->s1, xs);
  return fma(s1, x5, s);
}
static inline double sc_cos_poly(double x2, const sc_tab *p) {
  double x4 = x2 * x2;
  double c1 = fma(x2, p->c1, p->c0);
  double c2 = fma(x2, p->c4, p->c3);
  double x6 = x2 * x4;
  double c = fma(x4, p->c2, c1);
  return fma(c2, x6, c);
}

static inline double sc_reduce_large(uint32_t xi, int *np) {
  const uint32_t *arr = &k_inv_pio4[(xi >> 26) & 15];
  int shift = (xi >> 23) & 7;
  uint64_t n, res0, res1, res2;
  xi = (xi & 0xffffff) | 0x800000;
  xi <<= shift;
  res0 = (uint32_t)(xi * arr[0]);
  res1 = (uint64_t)xi * arr[4];
  res2 = (uint64_t)xi * arr[8];
  res0 = (res2 >> 32) | (res0 << 32);
  res0 += res1;
  n = (res0 + (1ULL << 61)) >> 62;
  res0 -= n << 62;
  *np = (int)n;
  return (double)(int64_t)res0 * 0x1.921FB54442D18p-62;
}

void orc_sincosf(float y, float *sinp, float *cosp) {
  double x = (double)y, x2, xs, rs, rc;
  uint32_t bits = f2u(y);
  uint32_t top = (bits >> 20) & 0x7ff;
  const sc_tab *p = &k_sc[0];
  int n = 0;
  if (top < 0x3f4) { /* |y| < pi/4 */
    if (top < 0x398) { /* |y| < 2^-12 */
      *sinp = y;
      *cosp = 1.0f;
      return;
    }
    x2 = x * x;
    xs = x;
  } else if (top < 0x42f) { /* |y| < 120 */
    double r = x * p->hpi_inv;
    n = ((int32_t)r + 0x800000) >> 24;
    x = fma(-(double)n, p->hpi, x);
    xs = x * p->sign[n & 3];
    if (n & 2) p = &k_sc[1];
    x2 = x * x;
  } else if (top < 0x7f8) {
    int sign = (int)(bits >> 31);
    x = sc_reduce_large(bits, &n);
    xs = x * p->sign[(n + sign) & 3];
    if ((n + sign) & 2) p = &k_sc[1];
    x2 = x * x;
  } else {
    *sinp = *cosp = y - y;
    return;
  }
  rs = sc_sin_poly(xs, x2, p);
  rc = sc_cos_poly(x2, p);
  if (n & 1) {
    *sinp = (float)rc;
    *cosp = (float)rs;
  } else {
    *sinp = (float)rs;
    *cosp = (float)rc;
  }
}

float orc_cosf(float y) {
  float s, c;
  orc_sincosf(y, &s, &c);
  return c;
}

/* Batch forms so the tests can sweep millions of arguments quickly. */
void orc_atan2f_batch(const float *y, const float *x, size_t n, float *out) {
  for (size_t i = 0; i < n; i++) out[i] = orc_atan2f(y[i], x[i]);
}
void orc_sincosf_batch(const float *x, size_t n, float *s, float *c) {
  for (size_t i = 0; i < n; i++) orc_sincosf(x[i], &s[i], &c[i]);
}

/* ======================================================================== */
/* Filter design.                                                           */
/* ======================================================================== */

/* filter.cpp:103-114 -- Hann-windowed sinc; centre index uses integer
 * division; the arithmetic is double with one rounding to float per store. */
void orc_lpf_design(float Fs, float Fc, unsigned short ntaps, float *h) {
  float norm_fc = Fc / (Fs / 2);
  int centre = (ntaps - 1) / 2;
  for (int i = 0; i < ntaps; i++) {
    float v;
    if (i == centre) {
      v = norm_fc;
    } else {
      double arg = ORC_PI * norm_fc * (i - centre);
      v = (float)(norm_fc * (sin(arg) / arg));
    }
    double w = sin(i * ORC_PI / ntaps);
    h[i] = (float)(v * (w * w));
  }
}

/* filter.cpp:83-99 -- band-pass: sinc(pass/2) * cos(i*pi*centre) * Hann. */
void orc_bpf_design(float Fs, float Fb, float Fe, unsigned short ntaps, float *h) {
  float norm_centre = ((Fe + Fb) / 2) / (Fs / 2);
  float norm_pass = (Fe - Fb) / (Fs / 2);
  int centre = (ntaps - 1) / 2;
  for (int i = 0; i < ntaps; i++) {
    float v;
    if (i == centre) {
      v = norm_pass;
    } else {
      double arg = ORC_PI * norm_pass / 2 * (i - centre);
      v = (float)(norm_pass * (sin(arg) / arg));
    }
    v = (float)(v * cos(i * ORC_PI * norm_centre));
    double w = sin(i * ORC_PI / ntaps);
    h[i] = (float)(v * w * w);
  }
}

/* ======================================================================== */
/* Primitives.                                                              */
/* ======================================================================== */

/* iofunc.cpp:128-135 -- (u8-128)/128.0, exact in float. */
void orc_u8_to_f32(const uint8_t *raw, size_t n, float *out) {
  for (size_t k = 0; k < n; k++) out[k] = (float)(((int)raw[k] - 128) / 128.0);
}

/* One output of a stateful FIR: taps ascending, multiply then add, float
 * accumulator starting from +0 (filter.cpp:137-146 / :163-178). */
static inline float fir_dot(const float *x, ptrdiff_t m, const float *h, size_t nh,
                            const float *state, size_t ns) {
  float acc = 0.0f;
  for (size_t n = 0; n < nh; n++) {
    ptrdiff_t idx = m - (ptrdiff_t)n;
    float xv = (idx >= 0) ? x[idx] : state[idx + (ptrdiff_t)ns];
    acc = acc + h[n] * xv;
  }
  return acc;
}

static void save_tail(float *state, size_t ns, const float *x, size_t nx) {
  /* filter.cpp:148-153 / :183-187 */
  for (size_t k = 0; k < ns; k++) state[k] = x[nx - ns + k];
}

void orc_fir_block(float *y, const float *x, size_t nx, const float *h, size_t nh,
                   float *state) {
  size_t ns = nh - 1;
  for (size_t n = 0; n < nx; n++) y[n] = fir_dot(x, (ptrdiff_t)n, h, nh, state, ns);
  save_tail(state, ns, x, nx);
}

void orc_fir_decim(float *y, const float *x, size_t nx, const float *h, size_t nh,
                   float *state, unsigned decim) {
  size_t ns = nh - 1;
  size_t ny = nx / decim;
  /* The reference iterates m <= nx (filter.cpp:166) and so performs one
   * out-of-bounds extra iteration; in-range outputs are unaffected and the
   * extra one is deliberately not restated. */
  for (size_t j = 0; j < ny; j++) y[j] = fir_dot(x, (ptrdiff_t)(j * decim), h, nh, state, ns);
  save_tail(state, ns, x, nx);
}

/* filter.cpp:191-223.  `state` keeps the reference's layout: nh-1 slots of the
 * zero-stuffed (upsampled) history, only slots U-1, 2U-1, ... are ever live. */
void orc_fir_resample(float *y, const float *x, size_t nx, const float *h, size_t nh,
                      float *state, unsigned decim, unsigned upsamp) {
  long ns = (long)nh - 1;
  long U = (long)upsamp, D = (long)decim;
  long total = (long)nx * U;
  for (long m = 0; m < total; m += D) {
    long phase = m % U;
    float acc = 0.0f;
    for (long n = phase; n < (long)nh; n += U) {
      long d = m - n;
      float xv = (d >= 0) ? x[d / U] : state[d + ns];
      acc = acc + h[n] * xv;
    }
    /* filter.cpp:213: y += y*U, i.e. gain (1+U) with two roundings */
    acc = acc + acc * (float)upsamp;
    y[m / D] = acc;
  }
  /* filter.cpp:217-222 */
  long k = U - 1;
  for (long i = U * (long)nx - ns; i < U * (long)nx - U; i += U) {
    state[k] = x[i / U + 1];
    k += U;
  }
}

/* filter.cpp:248-266 */
void orc_fm_demod(float *out, const float *I, const float *Q, size_t n, float *prev_i,
                  float *prev_q) {
  float pi = *prev_i, pq = *prev_q;
  for (size_t k = 0; k < n; k++) {
    float i = I[k], q = Q[k];
    float den = i * i + q * q;
    if (den == 0)
      out[k] = 0;
    else
      out[k] = (i * (q - pq) - q * (i - pi)) / den;
    pi = i;
    pq = q;
  }
  if (n) {
    *prev_i = I[n - 1];
    *prev_q = Q[n - 1];
  }
}

/* filter.cpp:14-29 -- a pure delay of ns samples. */
void orc_allpass(const float *in, size_t n, float *state, size_t ns, float *out) {
  for (size_t k = 0; k < ns; k++) out[k] = state[k];
  for (size_t k = ns; k < n; k++) out[k] = in[k - ns];
  for (size_t k = 0; k < ns; k++) state[k] = in[n - ns + k];
}

/* filter.cpp:32-80 */
void orc_pll(const float *in, size_t n, float *out, float *state, float freq, float Fs,
             float ncoScale, float phaseAdjust, float normBandwidth) {
  const float Cp = 2.666f, Ci = 3.555f;
  float Kp = normBandwidth * Cp;
  float Ki = (normBandwidth * normBandwidth) * Ci;
  float integrator = state[0], phaseEst = state[1];
  float feedbackI = state[2], feedbackQ = state[3];
  float trigOffset = state[5];
  float ratio = freq / Fs;
  out[0] = state[4];
  for (size_t k = 0; k < n; k++) {
    float errorI = in[k] * feedbackI;
    float errorQ = in[k] * (-feedbackQ);
    float errorD = orc_atan2f(errorQ, errorI);
    integrator = integrator + Ki * errorD;
    phaseEst = (phaseEst + Kp * errorD) + integrator;
    trigOffset += 1;
    /* filter.cpp:68: double expression stored into a float */
    float trigArg = (float)(((2 * ORC_PI) * (double)ratio) * (double)trigOffset + (double)phaseEst);
    orc_sincosf(trigArg, &feedbackQ, &feedbackI);
    out[k + 1] = orc_cosf(trigArg * ncoScale + phaseAdjust);
  }
  state[0] = integrator;
  state[1] = phaseEst;
  state[2] = feedbackI;
  state[3] = feedbackQ;
  state[4] = out[n];
  state[5] = trigOffset;
}

/* threadMonoOnly.cpp:185-190: NaN -> 0, else static_cast<short>(v*16384).  The
 * cast is cvttss2si (32-bit) followed by a 16-bit truncation on x86-64;
 * out-of-int32-range inputs give the "integer indefinite" 0x80000000. */
int16_t orc_pcm16(float v) {
  if (isnan(v)) return 0;
  float t = v * 16384;
  int32_t i;
  if (!(t >= -2147483648.0f && t < 2147483648.0f))
    i = INT32_MIN;
  else
    i = (int32_t)t;
  return (int16_t)(uint16_t)((uint32_t)i & 0xffffu);
}

/* ======================================================================== */
/* Whole chain.                                                             */
/* ======================================================================== */

int orc_mode_lookup(int mode, orc_mode_info *o) {
  /* project.cpp:424-427 mode table; :55-57 block sizes */
  switch (mode) {
    case 0:
      *o = (orc_mode_info){2400000, 240000, 48000, 10, 5, 1, 1024 * 10 * 5 * 2};
      return 0;
    case 1:
      *o = (orc_mode_info){1440000, 288000, 48000, 5, 6, 1, 1024 * 5 * 6 * 2};
      return 0;
    case 2:
      *o = (orc_mode_info){2400000, 240000, 44100, 10, 800, 147, 7 * 800 * 10 * 2};
      return 0;
    case 3:
      *o = (orc_mode_info){960000, 320000, 44100, 3, 3200, 441, 7 * 3200 * 3 * 2};
      return 0;
    default:
      return -1;
  }
}

typedef struct {
  float *p;
  size_t n, cap;
} fvec;

static void fvec_append(fvec *v, const float *src, size_t n) {
  if (v->n + n > v->cap) {
    size_t nc = v->cap ? v->cap * 2 : 4096;
    while (nc < v->n + n) nc *= 2;
    v->p = (float *)realloc(v->p, nc * sizeof(float));
    v->cap = nc;
  }
  memcpy(v->p + v->n, src, n * sizeof(float));
  v->n += n;
}

struct orc_chain {
  orc_config cfg;
  orc_mode_info mi;
  int resample;
  size_t n_audio_taps; /* audio_taps * U in modes 2/3 */
  float *rf_h, *audio_h, *pilot_h, *stereo_h;
  float *I_state, *Q_state, prev_i, prev_q;
  float *state_mono, *state_stereo, *state_carrier, *state_stereofilt, *state_allpass;
  float state_pll[6];
  /* per-block scratch */
  float *iq, *I, *Q, *I_filt, *Q_filt, *demod, *allp, *st_filt, *car_filt, *nco, *mixer;
  float *audio_filt, *stereo_final;
  size_t n_rf, n_if, n_audio;
  fvec taps[ORC_TAP_COUNT];
};

static float *zalloc(size_t n) { return (float *)calloc(n ? n : 1, sizeof(float)); }

void orc_chain_reset(orc_chain *c) {
  const orc_config *g = &c->cfg;
  memset(c->I_state, 0, sizeof(float) * (g->rf_taps - 1));
  memset(c->Q_state, 0, sizeof(float) * (g->rf_taps - 1));
  c->prev_i = c->prev_q = 0.0f;
  memset(c->state_mono, 0, sizeof(float) * (c->n_audio_taps - 1));
  memset(c->state_stereofilt, 0, sizeof(float) * (c->n_audio_taps - 1));
  memset(c->state_stereo, 0, sizeof(float) * (g->stereo_taps - 1));
  memset(c->state_carrier, 0, sizeof(float) * (g->stereo_taps - 1));
  memset(c->state_allpass, 0, sizeof(float) * ((g->stereo_taps - 1) / 2));
  /* project.cpp:458 */
  const float init[6] = {0.0f, 0.0f, 1.0f, 0.0f, 1.0f, 0.0f};
  memcpy(c->state_pll, init, sizeof init);
  orc_chain_clear_taps(c);
}

orc_chain *orc_chain_create(const orc_config *cfg) {
  orc_mode_info mi;
  if (orc_mode_lookup(cfg->mode, &mi)) return NULL;
  if (cfg->channels < 1 || cfg->channels > 2) return NULL;
  if (cfg->rf_taps < 2 || cfg->audio_taps < 2 || cfg->stereo_taps < 3) return NULL;
  orc_chain *c = (orc_chain *)calloc(1, sizeof *c);
  c->cfg = *cfg;
  c->mi = mi;
  c->resample = (cfg->mode >= 2);
  c->n_audio_taps = (size_t)cfg->audio_taps * (size_t)mi.audio_upsamp;
  if (c->n_audio_taps > 65535) { /* unsigned short in filter.h:24 */
    free(c);
    return NULL;
  }
  c->rf_h = zalloc(cfg->rf_taps);
  c->audio_h = zalloc(c->n_audio_taps);
  c->pilot_h = zalloc(cfg->stereo_taps);
  c->stereo_h = zalloc(cfg->stereo_taps);
  /* project.cpp:50 */
  orc_lpf_design((float)mi.rf_Fs, (float)100000, (unsigned short)cfg->rf_taps, c->rf_h);
  /* project.cpp:165-167 / :321-323 */
  orc_lpf_design((float)(mi.if_Fs * mi.audio_upsamp), (float)16000,
                 (unsigned short)c->n_audio_taps, c->audio_h);
  /* project.cpp:172-173 */
  orc_bpf_design((float)mi.if_Fs, (float)18.5e3, (float)19.5e3, (unsigned short)cfg->stereo_taps,
                 c->pilot_h);
  orc_bpf_design((float)mi.if_Fs, (float)22e3, (float)54e3, (unsigned short)cfg->stereo_taps,
                 c->stereo_h);
  c->I_state = zalloc(cfg->rf_taps - 1);
  c->Q_state = zalloc(cfg->rf_taps - 1);
  c->state_mono = zalloc(c->n_audio_taps - 1);
  c->state_stereofilt = zalloc(c->n_audio_taps - 1);
  c->state_stereo = zalloc(cfg->stereo_taps - 1);
  c->state_carrier = zalloc(cfg->stereo_taps - 1);
  c->state_allpass = zalloc((cfg->stereo_taps - 1) / 2);
  c->n_rf = (size_t)mi.block_bytes / 2;
  c->n_if = c->n_rf / mi.rf_decim;
  c->n_audio = c->n_if * mi.audio_upsamp / mi.audio_decim;
  c->iq = zalloc(mi.block_bytes);
  c->I = zalloc(c->n_rf);
  c->Q = zalloc(c->n_rf);
  c->I_filt = zalloc(c->n_if);
  c->Q_filt = zalloc(c->n_if);
  c->demod = zalloc(c->n_if);
  c->allp = zalloc(c->n_if);
  c->st_filt = zalloc(c->n_if);
  c->car_filt = zalloc(c->n_if);
  c->nco = zalloc(c->n_if + 1);
  c->mixer = zalloc(c->n_if);
  c->audio_filt = zalloc(c->n_audio);
  c->stereo_final = zalloc(c->n_audio);
  orc_chain_reset(c);
  return c;
}

void orc_chain_clear_taps(orc_chain *c) {
  for (int i = 0; i < ORC_TAP_COUNT; i++) c->taps[i].n = 0;
}

void orc_chain_destroy(orc_chain *c) {
  if (!c) return;
  float *all[] = {c->rf_h,        c->audio_h,       c->pilot_h,       c->stereo_h,
                  c->I_state,     c->Q_state,       c->state_mono,    c->state_stereofilt,
                  c->state_stereo, c->state_carrier, c->state_allpass, c->iq,
                  c->I,           c->Q,             c->I_filt,        c->Q_filt,
                  c->demod,       c->allp,          c->st_filt,       c->car_filt,
                  c->nco,         c->mixer,         c->audio_filt,    c->stereo_final};
  for (size_t i = 0; i < sizeof all / sizeof all[0]; i++) free(all[i]);
  for (int i = 0; i < ORC_TAP_COUNT; i++) free(c->taps[i].p);
  free(c);
}

const float *orc_chain_tap(const orc_chain *c, int stage, size_t *n) {
  if (stage < 0 || stage >= ORC_TAP_COUNT) {
    *n = 0;
    return NULL;
  }
  *n = c->taps[stage].n;
  return c->taps[stage].p;
}

static void audio_stage(orc_chain *c, float *y, const float *x, float *state) {
  const orc_mode_info *mi = &c->mi;
  if (!c->resample)
    orc_fir_decim(y, x, c->n_if, c->audio_h, c->n_audio_taps, state, (unsigned)mi->audio_decim);
  else
    orc_fir_resample(y, x, c->n_if, c->audio_h, c->n_audio_taps, state,
                     (unsigned)mi->audio_decim, (unsigned)mi->audio_upsamp);
}

size_t orc_chain_process(orc_chain *c, const uint8_t *iq, size_t nbytes, int16_t *pcm,
                         int keep) {
  const orc_config *g = &c->cfg;
  const orc_mode_info *mi = &c->mi;
  size_t nblocks = nbytes / (size_t)mi->block_bytes;
  size_t out = 0;
  for (size_t b = 0; b < nblocks; b++) {
    const uint8_t *blk = iq + b * (size_t)mi->block_bytes;
    /* project.cpp:82,101-105: normalise and de-interleave */
    orc_u8_to_f32(blk, (size_t)mi->block_bytes, c->iq);
    for (size_t k = 0; k < c->n_rf; k++) {
      c->I[k] = c->iq[2 * k];
      c->Q[k] = c->iq[2 * k + 1];
    }
    /* project.cpp:111,121,128 */
    orc_fir_decim(c->I_filt, c->I, c->n_rf, c->rf_h, (size_t)g->rf_taps, c->I_state,
                  (unsigned)mi->rf_decim);
    orc_fir_decim(c->Q_filt, c->Q, c->n_rf, c->rf_h, (size_t)g->rf_taps, c->Q_state,
                  (unsigned)mi->rf_decim);
    orc_fm_demod(c->demod, c->I_filt, c->Q_filt, c->n_if, &c->prev_i, &c->prev_q);
    if (keep) {
      fvec_append(&c->taps[ORC_TAP_I_FILT], c->I_filt, c->n_if);
      fvec_append(&c->taps[ORC_TAP_Q_FILT], c->Q_filt, c->n_if);
      fvec_append(&c->taps[ORC_TAP_DEMOD], c->demod, c->n_if);
    }
    if (g->channels == 1) {
      /* project.cpp:343-357 */
      audio_stage(c, c->audio_filt, c->demod, c->state_mono);
      if (keep) fvec_append(&c->taps[ORC_TAP_AUDIO_FILT], c->audio_filt, c->n_audio);
      for (size_t k = 0; k < c->n_audio; k++) pcm[out++] = orc_pcm16(c->audio_filt[k]);
    } else {
      /* project.cpp:191-280 */
      orc_allpass(c->demod, c->n_if, c->state_allpass, (size_t)(g->stereo_taps - 1) / 2, c->allp);
      orc_fir_block(c->st_filt, c->demod, c->n_if, c->stereo_h, (size_t)g->stereo_taps,
                    c->state_stereo);
      orc_fir_block(c->car_filt, c->demod, c->n_if, c->pilot_h, (size_t)g->stereo_taps,
                    c->state_carrier);
      audio_stage(c, c->audio_filt, c->allp, c->state_mono);
      orc_pll(c->car_filt, c->n_if, c->nco, c->state_pll, (float)19e3, (float)mi->if_Fs,
              (float)2.0, (float)0.0, (float)0.01);
      for (size_t z = 0; z < c->n_if; z++) c->mixer[z] = c->st_filt[z] * c->nco[z] * 2;
      audio_stage(c, c->stereo_final, c->mixer, c->state_stereofilt);
      if (keep) {
        fvec_append(&c->taps[ORC_TAP_ALLPASS], c->allp, c->n_if);
        fvec_append(&c->taps[ORC_TAP_STEREO_FILT], c->st_filt, c->n_if);
        fvec_append(&c->taps[ORC_TAP_CARRIER_FILT], c->car_filt, c->n_if);
        fvec_append(&c->taps[ORC_TAP_NCO], c->nco, c->n_if); /* [0..N): what the mixer sees */
        fvec_append(&c->taps[ORC_TAP_MIXER], c->mixer, c->n_if);
        fvec_append(&c->taps[ORC_TAP_AUDIO_FILT], c->audio_filt, c->n_audio);
        fvec_append(&c->taps[ORC_TAP_STEREO_FINAL], c->stereo_final, c->n_audio);
      }
      /* project.cpp:277-280 combine, :294-301 interleaved PCM */
      for (size_t s = 0; s < c->n_audio; s++) {
        float L = c->stereo_final[s] + c->audio_filt[s];
        float R = c->audio_filt[s] - c->stereo_final[s];
        pcm[out++] = orc_pcm16(L);
        pcm[out++] = orc_pcm16(R);
      }
    }
  }
  return out;
}
