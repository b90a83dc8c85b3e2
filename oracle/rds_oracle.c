/*
 * rds_oracle.c -- see rds_oracle.h.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain C, double precision, every multiply and add rounded separately
 * (-ffp-contract=off) like CPython / numpy scalar arithmetic.
 */
#include "rds_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------ */
/* design                                                                    */
/* ------------------------------------------------------------------------ */

/* fmSupportLib.py:358-372 */
void rdo_bandpass(int ntaps, double Fs, double Fb, double Fe, double *h) {
  const double Normcenter = ((Fe + Fb) / 2) / (Fs / 2);
  const double Normpass = (Fe - Fb) / (Fs / 2);
  const double mid = (double)(ntaps - 1) / 2; /* Python true division */
  for (int i = 0; i < ntaps; ++i) {
    double c;
    if ((double)i == mid) {
      c = Normpass;
    } else {
      const double a = M_PI * Normpass / 2 * ((double)i - mid);
      c = Normpass * (sin(a) / a);
    }
    c = c * cos((double)i * M_PI * Normcenter);
    const double w = sin((double)i * M_PI / (double)ntaps);
    c = c * (w * w);
    h[i] = c;
  }
}

/* fmSupportLib.py:376-385 */
void rdo_lowpass(int ntaps, double Fs, double Fc, double *h) {
  const double NormFc = Fc / (Fs / 2);
  const double mid = (double)(ntaps - 1) / 2;
  for (int i = 0; i < ntaps; ++i) {
    double c;
    if ((double)i == mid) {
      c = NormFc;
    } else {
      const double a = M_PI * NormFc * ((double)i - mid);
      c = NormFc * (sin(a) / a);
    }
    const double w = sin((double)i * M_PI / (double)ntaps);
    h[i] = c * (w * w);
  }
}

/* fmSupportLib.py:251-287 */
void rdo_rrc(double Fs, int ntaps, double *h) {
  const double T_symbol = 1 / 2375.0;
  const double beta = 0.90;
  for (int k = 0; k < ntaps; ++k) {
    const double t = ((double)k - (double)ntaps / 2) / Fs; /* N_taps/2 is a true division */
    if (t == 0.0) {
      h[k] = 1.0 + beta * ((4 / M_PI) - 1);
    } else if (t == -T_symbol / (4 * beta) || t == T_symbol / (4 * beta)) {
      h[k] = (beta / sqrt(2.0)) * (((1 + 2 / M_PI) * (sin(M_PI / (4 * beta)))) +
                                   ((1 - 2 / M_PI) * (cos(M_PI / (4 * beta)))));
    } else {
      const double q = 4 * beta * t / T_symbol;
      h[k] = (sin(M_PI * t * (1 - beta) / T_symbol) +
              4 * beta * (t / T_symbol) * cos(M_PI * t * (1 + beta) / T_symbol)) /
             (M_PI * t * (1 - q * q) / T_symbol);
    }
  }
}

/* ------------------------------------------------------------------------ */
/* streaming primitives                                                      */
/* ------------------------------------------------------------------------ */

/* lfilter(h, 1.0, x, zi): y[n] = sum_k h[k] x[n-k] across block boundaries
 * (fmRDS.py:223,233,248). */
void rdo_fir(const double *x, size_t n, const double *h, int nh, double *hist, double *y) {
  const int S = nh - 1;
  for (size_t i = 0; i < n; ++i) {
    double acc = 0.0;
    for (int k = 0; k < nh; ++k) {
      const long j = (long)i - k;
      const double v = j >= 0 ? x[j] : hist[S + j];
      acc += h[k] * v;
    }
    y[i] = acc;
  }
  /* next history = last S samples of (hist ++ x) */
  if (n >= (size_t)S) {
    memcpy(hist, x + n - S, (size_t)S * sizeof(double));
  } else {
    memmove(hist, hist + n, ((size_t)S - n) * sizeof(double));
    memcpy(hist + S - n, x, n * sizeof(double));
  }
}

/* fmSupportLib.py:291-295 */
void rdo_allpass(const double *x, size_t n, double *state, int ns, double *y) {
  double *tmp = (double *)malloc((size_t)ns * sizeof(double));
  memcpy(tmp, x + n - ns, (size_t)ns * sizeof(double));
  for (size_t i = n; i-- > (size_t)ns;) y[i] = x[i - ns];
  memcpy(y, state, (size_t)ns * sizeof(double));
  memcpy(state, tmp, (size_t)ns * sizeof(double));
  free(tmp);
}

/* fmSupportLib.py:297-354 */
void rdo_pll(const double *x, size_t n, double freq, double Fs, double *state, double ncoScale,
             double phaseAdjust, double normBandwidth, double *outI, double *outQ) {
  const double Cp = 2.666, Ci = 3.555;
  const double Kp = normBandwidth * Cp;
  const double Ki = (normBandwidth * normBandwidth) * Ci;
  double integrator = state[0], phaseEst = state[1], feedbackI = state[2], feedbackQ = state[3];
  double trigOffset = state[5];
  outI[0] = state[4];
  outQ[0] = state[6];
  for (size_t k = 0; k < n; ++k) {
    const double errorI = x[k] * (+feedbackI);
    const double errorQ = x[k] * (-feedbackQ);
    const double errorD = atan2(errorQ, errorI);
    integrator = integrator + Ki * errorD;
    phaseEst = phaseEst + Kp * errorD + integrator;
    trigOffset += 1;
    const double trigArg = 2 * M_PI * (freq / Fs) * trigOffset + phaseEst;
    feedbackI = cos(trigArg);
    feedbackQ = sin(trigArg);
    outI[k + 1] = cos(trigArg * ncoScale + phaseAdjust);
    outQ[k + 1] = sin(trigArg * ncoScale + phaseAdjust);
  }
  state[0] = integrator;
  state[1] = phaseEst;
  state[2] = feedbackI;
  state[3] = feedbackQ;
  state[4] = outI[n];
  state[5] = trigOffset;
  state[6] = outQ[n];
}

/* fmSupportLib.py:388-407.  The model keeps a zero-stuffed state of nh-1 slots whose live
 * entries state[kU-1] (k = 1..nh/U-1) are the last nh/U-1 inputs (:402-405); m-n is always a
 * multiple of U, so only live slots are ever read (:396-399).  hist is that compact form:
 * hist[i] = x_prev[N - (nh/U-1) + i]. */
void rdo_resample(const double *x, size_t n, const double *h, int nh, double *hist, int decim,
                  int upsamp, double *y) {
  const int U = upsamp, D = decim;
  const int TP = nh / U; /* taps per phase */
  const int S = TP - 1;
  const size_t n_out = n * (size_t)U / (size_t)D;
  for (size_t j = 0; j < n_out; ++j) {
    const long long m = (long long)j * D;
    const int phase = (int)(m % U);
    const long long base = (m - phase) / U;
    double acc = 0.0;
    for (int k = 0; k < TP; ++k) {
      const long long idx = base - k;
      const double v = idx >= 0 ? x[idx] : hist[S + idx];
      acc += h[phase + (long long)k * U] * v;
    }
    y[j] = acc * (double)U;
  }
  if (n >= (size_t)S) {
    memcpy(hist, x + n - S, (size_t)S * sizeof(double));
  } else {
    memmove(hist, hist + n, ((size_t)S - n) * sizeof(double));
    memcpy(hist + S - n, x, n * sizeof(double));
  }
}

/* ------------------------------------------------------------------------ */
/* bit layer                                                                 */
/* ------------------------------------------------------------------------ */

/* fmSupportLib.py:228-236: only pair[0] is looked at */
static int symbol_to_bit(const double *pair) {
  if (pair[0] < 0) return 0;
  if (pair[0] > 0) return 1;
  return 0;
}

/* fmSupportLib.py:103-201 with an explicit to_pass_on_state = {pair[0], pair[1], start,
 * prev_size} (:104-106), updated on return as at :178-189.  fmRDS.py:257-260 re-creates the
 * state {0, 0, 158, 0} for every block (rdo_cdr below); carried from block to block it is
 * what the function was written for: an odd number of sampling points leaves one symbol that
 * is paired with the first point of the next block (:117-125). */
int rdo_cdr_state(const double *x, int n, int sps, int block_count, double *state, uint8_t *bits, int cap) {
  double pair[2] = {state[0], state[1]};
  const int start_init = (int)state[2];
  const int prev_size = (int)state[3];
  int start = start_init;
  const double limit = 0.3;
  int n_prefix = 0; /* bits appended by the pairing branch (:117-125) and by restarts (:160-166) */
  const int max_samples = n / sps + 3;
  double *spa = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  double *samples = (double *)calloc((size_t)max_samples, sizeof(double));
  int size = 0;
  int unpaired = 1;
  while (unpaired) {
    memset(spa, 0, (size_t)n * sizeof(double));
    size = 0;
    const int loop_start = start;
    for (int i = loop_start; i < n; i += sps) {
      /* :117-125 the last symbol of the previous block meets the first of this one */
      if (i == start && start == start_init && (prev_size % 2 == 1)) {
        pair[1] = x[i];
        if (n_prefix < cap) bits[n_prefix] = (uint8_t)symbol_to_bit(pair);
        n_prefix++;
        pair[0] = pair[1];
        start = start + sps;
        continue;
      }
      /* :128-136: a third consecutive high (or low) is inverted */
      if (i >= start + 2 * sps && spa[i - 2 * sps] > 0 && spa[i - sps] > 0 && x[i] > 0)
        spa[i] = -1 * x[i];
      else if (i >= start + 2 * sps && spa[i - 2 * sps] < 0 && spa[i - sps] < 0 && x[i] < 0)
        spa[i] = -1 * x[i];
      else
        spa[i] = x[i];
      size += 1;
    }
    /* :143 copyFrom */
    memset(samples, 0, (size_t)max_samples * sizeof(double));
    for (int i = start; i < n; i += sps) samples[(i - start) / sps] = spa[i];
    int restarted = 0, any_good = 0;
    for (int i = 0; i < size; i += 2) {
      if (i + 1 < size) {
        if ((samples[i] < 0 && samples[i + 1] < 0) || (samples[i] > 0 && samples[i + 1] > 0)) {
          if (fabs(samples[i]) < limit || fabs(samples[i + 1]) < limit) {
            if (fabs(samples[i]) < limit) samples[i] = -1 * samples[i];
            else if (fabs(samples[i + 1]) < limit) samples[i + 1] = -1 * samples[i + 1];
          } else {
            start = start + sps;
            if (block_count != 0) {
              pair[1] = samples[0];
              if (n_prefix < cap) bits[n_prefix] = (uint8_t)symbol_to_bit(pair);
              n_prefix++;
              pair[0] = pair[1];
            }
            restarted = 1;
            break;
          }
        } else {
          any_good = 1;
        }
      }
    }
    /* :170-176: a restart runs the pass again from the new start.  Otherwise the reference
     * leaves the loop when some pair had opposite signs -- and never leaves it when none had
     * (e.g. fewer than two samples left after many restarts); here that pass is accepted. */
    (void)any_good;
    unpaired = restarted;
  }
  /* :178-189 state for the next block (with no sampling point left the reference would fail
   * on samples[-1]; the symbol in hand is kept) */
  if (size > 0) pair[0] = samples[size - 1];
  state[0] = pair[0];
  state[1] = pair[1];
  {
    const int last_index = ((size - 1) * sps) + start;
    state[2] = (double)(sps - (n - last_index));
  }
  state[3] = (double)size;
  /* :192 manchestering (fmSupportLib.py:203-222) */
  int nb = n_prefix;
  for (int i = 0; i < size; i += 2) {
    if (i + 1 < size) {
      uint8_t b = 0;
      if (samples[i] < 0 && samples[i + 1] > 0) b = 0;
      else if (samples[i] > 0 && samples[i + 1] < 0) b = 1;
      if (nb < cap) bits[nb] = b;
      nb++;
    }
  }
  free(spa);
  free(samples);
  return nb;
}

/* fmSupportLib.py:103-201 driven as at fmRDS.py:257-268: fresh state for every block */
int rdo_cdr(const double *x, int n, int sps, int block_count, uint8_t *bits, int cap) {
  double state[4] = {0.0, 0.0, 158.0, 0.0}; /* fmRDS.py:257-260 */
  return rdo_cdr_state(x, n, sps, block_count, state, bits, cap);
}

/* fmSupportLib.py:241-249 */
void rdo_diff_decode(const uint8_t *in, int n, uint8_t *out) {
  if (n <= 0) return;
  out[0] = in[0];
  for (int i = 1; i < n; ++i) out[i] = (uint8_t)(in[i] != in[i - 1]);
}

/* fmSupportLib.py:32-57 */
static const uint8_t kParity[26][10] = {
    {1, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {0, 1, 0, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 1, 0, 0, 0, 0, 0, 0, 0},
    {0, 0, 0, 1, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 1, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 1, 0, 0, 0, 0},
    {0, 0, 0, 0, 0, 0, 1, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 1, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0, 1, 0},
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 1}, {1, 0, 1, 1, 0, 1, 1, 1, 0, 0}, {0, 1, 0, 1, 1, 0, 1, 1, 1, 0},
    {0, 0, 1, 0, 1, 1, 0, 1, 1, 1}, {1, 0, 1, 0, 0, 0, 0, 1, 1, 1}, {1, 1, 1, 0, 0, 1, 1, 1, 1, 1},
    {1, 1, 0, 0, 0, 1, 0, 0, 1, 1}, {1, 1, 0, 1, 0, 1, 0, 1, 0, 1}, {1, 1, 0, 1, 1, 1, 0, 1, 1, 0},
    {0, 1, 1, 0, 1, 1, 1, 0, 1, 1}, {1, 0, 0, 0, 0, 0, 0, 0, 0, 1}, {1, 1, 1, 1, 0, 1, 1, 1, 0, 0},
    {0, 1, 1, 1, 1, 0, 1, 1, 1, 0}, {0, 0, 1, 1, 1, 1, 0, 1, 1, 1}, {1, 0, 1, 0, 1, 0, 0, 1, 1, 1},
    {1, 1, 1, 0, 0, 0, 1, 1, 1, 1}, {1, 1, 0, 0, 0, 1, 1, 0, 1, 1}};

/* fmSupportLib.py:14-27 */
void rdo_syndrome(const uint8_t *d, uint8_t *s) {
  for (int k = 0; k < 10; ++k) {
    int ones = 0;
    for (int i = 0; i < 26; ++i)
      if (d[i] * kParity[i][k] == 1) ones++;
    s[k] = (uint8_t)(ones % 2);
  }
}

static char offset_of(const uint8_t *s) {
  static const uint8_t A[10] = {1, 1, 1, 1, 0, 1, 1, 0, 0, 0};
  static const uint8_t B[10] = {1, 1, 1, 1, 0, 1, 0, 1, 0, 0};
  static const uint8_t Cw[10] = {1, 0, 0, 1, 0, 1, 1, 1, 0, 0};
  static const uint8_t Cp[10] = {1, 1, 1, 1, 0, 0, 1, 1, 0, 0};
  static const uint8_t Dw[10] = {1, 0, 0, 1, 0, 1, 1, 0, 0, 0};
  if (!memcmp(s, A, 10)) return 'A';
  if (!memcmp(s, B, 10)) return 'B';
  if (!memcmp(s, Cw, 10)) return 'C';
  if (!memcmp(s, Cp, 10)) return 'c';
  if (!memcmp(s, Dw, 10)) return 'D';
  return ' ';
}

/* fmSupportLib.py:30-100 */
char rdo_framesync(const uint8_t *d, int n, int *state_index) {
  int pos = 0;
  char offset_type = ' ';
  while (pos < n - 26) {
    uint8_t s[10];
    rdo_syndrome(d + pos, s);
    const char o = offset_of(s);
    if (o != ' ') {
      offset_type = o;
      if (n - (pos + 26) < 26) break;
      pos += 26;
    } else {
      pos += 1;
    }
  }
  *state_index = (offset_type == ' ') ? pos : pos + 26;
  return offset_type;
}

/* ------------------------------------------------------------------------ */
/* chain (fmRDS.py:55-75,100-125,149-192,222-274)                            */
/* ------------------------------------------------------------------------ */
struct rdo_chain {
  int mode, n, n_out, U, D, sps;
  double if_Fs;
  double *h_chan, *h_carr, *h_rs, *h_rrc;
  int nh_rs;
  double *hist_chan, *hist_carr, *st_allpass, *hist_rsI, *hist_rsQ, *hist_rrcI, *hist_rrcQ;
  double pll[7];
  double *chan, *allp, *sq, *carr, *pllI, *pllQ, *mixI, *mixQ, *rsI, *rsQ, *rrcI, *rrcQ;
  uint8_t *cdr_bits, *diff_bits;
  int n_bits;
  uint8_t *decoded;
  int n_decoded, cap_decoded;
  char offset;
  int block_count;
  int cdr_carry;       /* keep the CDR state from block to block instead of fmRDS.py:257-260 */
  double cdr_state[4];
};

static double *dz(size_t n) { return (double *)calloc(n ? n : 1, sizeof(double)); }

rdo_chain *rdo_chain_create(int mode, int block_if) {
  if (mode != 0 && mode != 2) return NULL;
  rdo_chain *c = (rdo_chain *)calloc(1, sizeof(*c));
  c->mode = mode;
  c->if_Fs = 240000.0;
  c->U = mode == 0 ? 247 : 817;  /* fmRDS.py:57,69 */
  c->D = mode == 0 ? 960 : 1920; /* fmRDS.py:58,70 */
  c->sps = mode == 0 ? 26 : 43;  /* fmRDS.py:60,72 */
  c->n = block_if;
  c->n_out = (int)((long long)block_if * c->U / c->D);
  const int T = 151; /* fmRDS.py:100 */
  c->h_chan = dz(T);
  c->h_carr = dz(T);
  rdo_bandpass(T, c->if_Fs, 54e3, 60e3, c->h_chan);      /* fmRDS.py:122 */
  rdo_bandpass(T, c->if_Fs, 113.5e3, 114.5e3, c->h_carr); /* fmRDS.py:123 */
  c->nh_rs = 101 * c->U;                                  /* fmRDS.py:59,71 */
  c->h_rs = dz((size_t)c->nh_rs);
  rdo_lowpass(c->nh_rs, c->if_Fs * c->U, 3e3, c->h_rs); /* fmRDS.py:124 */
  c->h_rrc = dz(101);
  rdo_rrc(2375.0 * c->sps, 101, c->h_rrc); /* fmRDS.py:125 */
  c->hist_chan = dz(T - 1);
  c->hist_carr = dz(T - 1);
  c->st_allpass = dz((T - 1) / 2);
  c->hist_rsI = dz(100);
  c->hist_rsQ = dz(100);
  c->hist_rrcI = dz(100);
  c->hist_rrcQ = dz(100);
  const double init[7] = {0.0, 0.0, 1.0, 0.0, 1.0, 0, 1.0}; /* fmRDS.py:175 */
  memcpy(c->pll, init, sizeof(init));
  const size_t n = (size_t)c->n, no = (size_t)c->n_out;
  c->chan = dz(n); c->allp = dz(n); c->sq = dz(n); c->carr = dz(n);
  c->pllI = dz(n + 1); c->pllQ = dz(n + 1); c->mixI = dz(n); c->mixQ = dz(n);
  c->rsI = dz(no); c->rsQ = dz(no); c->rrcI = dz(no); c->rrcQ = dz(no);
  const int cap = c->n_out / c->sps + 4;
  c->cdr_bits = (uint8_t *)calloc((size_t)cap, 1);
  c->diff_bits = (uint8_t *)calloc((size_t)cap, 1);
  c->cap_decoded = 0;
  c->decoded = NULL;
  c->cdr_state[2] = 158.0; /* fmRDS.py:259 */
  return c;
}

void rdo_chain_destroy(rdo_chain *c) {
  if (!c) return;
  double *d[] = {c->h_chan, c->h_carr, c->h_rs, c->h_rrc, c->hist_chan, c->hist_carr, c->st_allpass,
                 c->hist_rsI, c->hist_rsQ, c->hist_rrcI, c->hist_rrcQ, c->chan, c->allp, c->sq,
                 c->carr, c->pllI, c->pllQ, c->mixI, c->mixQ, c->rsI, c->rsQ, c->rrcI, c->rrcQ};
  for (size_t i = 0; i < sizeof(d) / sizeof(d[0]); ++i) free(d[i]);
  free(c->cdr_bits);
  free(c->diff_bits);
  free(c->decoded);
  free(c);
}

void rdo_chain_block(rdo_chain *c, const double *fm_demod) {
  const size_t n = (size_t)c->n;
  rdo_fir(fm_demod, n, c->h_chan, 151, c->hist_chan, c->chan);                 /* :223 */
  rdo_allpass(c->chan, n, c->st_allpass, 75, c->allp);                          /* :227 */
  for (size_t i = 0; i < n; ++i) c->sq[i] = c->chan[i] * c->chan[i];            /* :230 */
  rdo_fir(c->sq, n, c->h_carr, 151, c->hist_carr, c->carr);                     /* :233 */
  rdo_pll(c->carr, n, 114e3, c->if_Fs, c->pll, 0.5, 3 * M_PI / 8, 0.002, c->pllI, c->pllQ); /* :236 */
  for (size_t i = 0; i < n; ++i) c->mixI[i] = c->pllI[i] * c->allp[i] * 2;      /* :241 */
  rdo_resample(c->mixI, n, c->h_rs, c->nh_rs, c->hist_rsI, c->D, c->U, c->rsI); /* :244 */
  rdo_fir(c->rsI, (size_t)c->n_out, c->h_rrc, 101, c->hist_rrcI, c->rrcI);      /* :248 */
  for (size_t i = 0; i < n; ++i) c->mixQ[i] = c->pllQ[i] * c->allp[i] * 2;      /* :251 */
  rdo_resample(c->mixQ, n, c->h_rs, c->nh_rs, c->hist_rsQ, c->D, c->U, c->rsQ); /* :252 */
  rdo_fir(c->rsQ, (size_t)c->n_out, c->h_rrc, 101, c->hist_rrcQ, c->rrcQ);      /* :254 */
  const int cap = c->n_out / c->sps + 4;
  if (c->cdr_carry)
    c->n_bits = rdo_cdr_state(c->rrcI, c->n_out, c->sps, c->block_count, c->cdr_state, c->cdr_bits, cap);
  else
    c->n_bits = rdo_cdr(c->rrcI, c->n_out, c->sps, c->block_count, c->cdr_bits, cap); /* :268 */
  rdo_diff_decode(c->cdr_bits, c->n_bits, c->diff_bits);                             /* :271 */
  if (c->n_decoded + c->n_bits > c->cap_decoded) {
    c->cap_decoded = 2 * (c->n_decoded + c->n_bits) + 64;
    c->decoded = (uint8_t *)realloc(c->decoded, (size_t)c->cap_decoded);
  }
  memcpy(c->decoded + c->n_decoded, c->diff_bits, (size_t)c->n_bits); /* :272 */
  c->n_decoded += c->n_bits;
  int used = 0;
  c->offset = rdo_framesync(c->decoded, c->n_decoded, &used); /* :275 */
  if (used > c->n_decoded) used = c->n_decoded;               /* Python slice semantics */
  memmove(c->decoded, c->decoded + used, (size_t)(c->n_decoded - used)); /* :276 */
  c->n_decoded -= used;
  c->block_count++;
}

void rdo_chain_set_cdr_carry(rdo_chain *c, int carry) { c->cdr_carry = carry; }

const double *rdo_chain_tap(const rdo_chain *c, int stage, size_t *n) {
  const size_t ni = (size_t)c->n, no = (size_t)c->n_out;
  switch (stage) {
    case 0: *n = ni; return c->chan;
    case 1: *n = ni; return c->carr;
    case 2: *n = ni + 1; return c->pllI;
    case 3: *n = ni + 1; return c->pllQ;
    case 4: *n = ni; return c->mixI;
    case 5: *n = ni; return c->mixQ;
    case 6: *n = no; return c->rsI;
    case 7: *n = no; return c->rsQ;
    case 8: *n = no; return c->rrcI;
    case 9: *n = no; return c->rrcQ;
  }
  *n = 0;
  return NULL;
}

int rdo_chain_bits(const rdo_chain *c, const uint8_t **cdr_bits, const uint8_t **diff_bits) {
  if (cdr_bits) *cdr_bits = c->cdr_bits;
  if (diff_bits) *diff_bits = c->diff_bits;
  return c->n_bits;
}

char rdo_chain_offset(const rdo_chain *c) { return c->offset; }
