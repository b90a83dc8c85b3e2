// rds.cu -- the RDS receiver chain on the device, attached to a batched pipeline.
//
// The reference has no C++ RDS path; what it has is the Python model
// model/fmRDS.py:222-276 on model/fmSupportLib.py, which
// works in double precision.  This file follows that model, stage for stage, in double
// precision, from the pipeline's fm_demod (float, bit-identical to the C++ reference):
//
//   R1  k_rds_fir<151, float in>     channel band-pass 54-60 kHz            fmRDS.py:223
//   R2  k_rds_fir<151, squared in>   x^2, carrier band-pass 113.5-114.5 kHz fmRDS.py:230-233
//   R3a k_rds_pll_warp / k_rds_pll   PLL at 114 kHz: a warp per capture solving 32 samples at a time
//                                    (up to 4096 captures), or one lane per capture  fmRDS.py:236
//   R3b k_rds_mix                    NCO I and Q (scale 0.5, 3pi/8), all-pass delay (75)
//                                    and both mixers                       fmRDS.py:227,241,251
//   R4  k_rds_resample               rational resampler U/D, 101 taps per phase, gain U
//                                                                           fmRDS.py:244,252
//   R5  k_rds_fir<101>               root-raised-cosine filter, I and Q     fmRDS.py:248,254
//   R6  k_rds_cdr                    clock/data recovery + Manchester decoding, one thread
//                                    per (capture, block)                   fmRDS.py:257-268
//   R7  k_rds_carry                  history prefixes for the next call
//   host: differential decoding and the frame synchroniser (fmRDS.py:271-276) on the few
//   bits per block that come back.
//
// Every row keeps the tail of the previous call in front of the current samples (the same
// "history prefix" layout as the audio path), so a call may cover any whole number of RDS
// blocks and the result does not depend on how a capture is cut into calls -- except for the
// CDR, whose window is the block, as in the model (its state is re-created per block,
// fmRDS.py:257-260).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/sdr_b200.h"
#include "common.cuh"
#include "pipeline_view.h"

using namespace sdr;

namespace sdr {

constexpr int RDS_T = 151;   // fmRDS.py:100 rds_taps
constexpr int RDS_TP = 101;  // taps per resampler phase (fmRDS.py:59,71) and RRC taps (:61,73)
constexpr int RDS_HC = 152;  // channel_filt history prefix: >= 150 (carrier FIR), >= 75 (all-pass)
constexpr int RDS_HM = 104;  // mixer history prefix: the resampler reaches 100 samples back
constexpr int RDS_HS = 104;  // resampler-output history prefix: the RRC reaches 100 back
constexpr int RDS_DELAY = (RDS_T - 1) / 2;  // fmRDS.py:178 state_rds_allpass
constexpr int RDS_CDR_START = 158;          // fmRDS.py:259 start_init

template <int T, typename E = double>
struct DTaps {
  E h[T];
};

// ---------------------------------------------------------------------------
// R1/R2/R5: double-precision FIR, y[n] = sum_k h[k] f(x[n-k]).
// One thread produces R consecutive outputs from a shared-memory window; the window is
// stored in groups of R with one padding slot so the per-thread stride (R+1 doubles) keeps
// the 64-bit loads of a half-warp on distinct bank pairs.  Taps are a __grid_constant__
// parameter: after unrolling each one is a constant-bank operand of its DFMA.
// ---------------------------------------------------------------------------
struct RdsFirArgs {
  const void *src[2];  // [B][src_stride]; z = 0/1 selects the row set (I / Q)
  size_t src_stride;
  int src_off;         // element of a row that holds sample 0 of this call (history before it)
  const float *hist32; // INMODE 0 only: [B][hist_len] samples that precede sample 0
  int hist_len;
  double *dst[2];
  size_t dst_stride;
  int dst_off;
  int n;
  int outs_per_seg;
};

// INMODE 0: float input with a separate history; 1: double input, squared; 2: double input.
// E = double is the model's arithmetic.  E = float (sdr_rds_config.precision = SDR_RDS_F32_FIR) keeps
// rows and taps' values but multiplies and accumulates in single precision: the experiment the
// survey's parity bar allows ("f32 GPU vs f64 model", RRC output <= 1e-5) -- see DESIGN.md 4b for
// what it buys and what it costs.
template <int T, int R, int NT, int INMODE, typename E = double>
__global__ void __launch_bounds__(NT)
k_rds_fir(const RdsFirArgs a, const __grid_constant__ DTaps<T, E> taps) {
  constexpr int HALO = ((T - 1 + R - 1) / R) * R;
  constexpr int TILE = NT * R;
  constexpr int WIN = HALO + TILE;
  __shared__ E xs[WIN + WIN / R + 1];
  const int t = threadIdx.x;
  const int b = blockIdx.y;
  const int z = blockIdx.z;
  const int o_begin = blockIdx.x * a.outs_per_seg;
  const int o_end = min(o_begin + a.outs_per_seg, a.n);
  const float *s32 = reinterpret_cast<const float *>(z ? a.src[1] : a.src[0]) + (size_t)b * a.src_stride + a.src_off;
  const double *s64 = reinterpret_cast<const double *>(z ? a.src[1] : a.src[0]) + (size_t)b * a.src_stride + a.src_off;
  const float *hist = INMODE == 0 ? a.hist32 + (size_t)b * a.hist_len : nullptr;
  double *drow = (z ? a.dst[1] : a.dst[0]) + (size_t)b * a.dst_stride + a.dst_off;
  for (int o0 = o_begin; o0 < o_end; o0 += TILE) {
    __syncthreads();
    for (int q = t; q < WIN; q += NT) {
      const int i = o0 - HALO + q;
      double v = 0.0;
      if (i < a.n) {
        if (INMODE == 0) {
          if (i >= 0) v = (double)s32[i];
          else if (i >= -a.hist_len) v = (double)hist[a.hist_len + i];
        } else if (i >= -a.src_off) {
          v = s64[i];
          if (INMODE == 1) v = __dmul_rn(v, v);  // fmRDS.py:230
        }
      }
      xs[q + q / R] = (E)v;
    }
    __syncthreads();
    E acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = (E)0;
    const E *w = xs + (t * R + HALO) + (t * R + HALO) / R;
#pragma unroll
    for (int c = R - 1; c >= -(T - 1); --c) {
      // window slot of sample (output 0 of this thread) + c, with the group padding
      const int off = c + (c >= 0 ? c / R : -((-c + R - 1) / R));
      const E v = w[off];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int k = r - c;
        if (k >= 0 && k < T) acc[r] = fma(taps.h[k], v, acc[r]);
      }
    }
    const int o = o0 + t * R;
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (o + r < o_end) drow[o + r] = (double)acc[r];
  }
}

// ---------------------------------------------------------------------------
// R3a: PLL (fmSupportLib.py:297-354).  Sequential per capture: one lane per capture, so what
// counts is the length of the dependent chain per sample.  The model's chain is
//     errorD = atan2(x * -sin(t), x * cos(t));  integrator += Ki*errorD;
//     phaseEst += Kp*errorD + integrator;       t' = 2*pi*(f/Fs)*n + phaseEst
// with t the NCO phase of the previous sample.  The phase detector's arguments are a real
// sample times a unit phasor, so its result is that phasor's angle: -t wrapped to (-pi, pi]
// when x > 0, pi - t wrapped when x < 0 (and atan2's signed-zero cases when x == 0).  The
// kernel therefore carries t reduced modulo 2*pi (two fused multiply-adds against a two-part
// 2*pi) instead of sin/cos of it, and no arctangent, sine or cosine is left on the chain:
// eight dependent double-precision operations (8 cycles each on B200) and one integer
// test-and-select per sample instead of ~100 operations, and no branch that depends on them.  The result
// differs from the model's only by the rounding of sin, cos and atan2 themselves (~1e-16 in
// errorD), far inside the parity bound.  The NCO phase t' of every sample goes to `theta`;
// R3b turns it into the NCO outputs and the mixer products in parallel.
// theta rows: [0] = last phase of the previous call (NaN before the first sample), [1+k] = t'
// of sample k.
// ---------------------------------------------------------------------------
struct RdsPllArgs {
  const double *carr;  // [B][carr_stride]
  size_t carr_stride;
  double *theta;  // [B][theta_stride]
  size_t theta_stride;
  double *state;  // [B][8]: integrator, phaseEst, reduced NCO phase, -, -, trigOffset
  int n, batch;
  double freq, Fs, normBandwidth;
};

constexpr int RDS_PLL_PITCH = 33;  // doubles per tile row: per-lane 64-bit reads fall on distinct bank pairs
constexpr double RDS_PI = 3.141592653589793;
constexpr double RDS_2PI_HI = 6.283185307179586;        // fl(2*pi)
constexpr double RDS_2PI_LO = 2.4492935982947064e-16;   // 2*pi - fl(2*pi)
constexpr double RDS_INV_2PI = 0.15915494309189535;

// x - 2*pi*rint(x / (2*pi)), |result| <= pi (+ a rounding), absolute error ~3e-16 for |x| < 2^40
__device__ __forceinline__ double rds_reduce_2pi(double x) {
  const double k = rint(__dmul_rn(x, RDS_INV_2PI));
  return fma(-k, RDS_2PI_LO, fma(-k, RDS_2PI_HI, x));
}

static __global__ void __launch_bounds__(32) k_rds_pll(const RdsPllArgs a) {
  const bool live = (int)(blockIdx.x * blockDim.x + threadIdx.x) < a.batch;
  const int b = live ? blockIdx.x * blockDim.x + threadIdx.x : a.batch - 1;  // idle lanes shadow the last capture
  // fmSupportLib.py:303-309
  const double Kp = __dmul_rn(a.normBandwidth, 2.666);
  const double Ki = __dmul_rn(__dmul_rn(a.normBandwidth, a.normBandwidth), 3.555);
  double *st = a.state + (size_t)b * 8;
  double integrator = st[0], phaseEst = st[1], r = st[2];
  const double n0 = st[5];
  double *th = a.theta + (size_t)b * a.theta_stride + 1;  // (shadow lanes store the same values as the lane they shadow)
  // fmSupportLib.py:340: 2*pi*(freq/Fs), left to right
  const double w = __dmul_rn(__dmul_rn(2.0, RDS_PI), __ddiv_rn(a.freq, a.Fs));
  // Input: the warp copies tiles of 32 captures x 32 samples into shared memory with
  // asynchronous copies (32 coalesced 256-byte rows per tile, no register in between), one
  // tile ahead of the one being consumed, so no global-load latency ever meets the recurrence.
  // (Per-lane loads into a register queue stall on the queue's register moves: 10.9 ms per
  // 76 800 samples against 6.4 ms with the loads removed.)
  __shared__ double tile[2][32][RDS_PLL_PITCH];
  const int lane = threadIdx.x;
  const int b_first = blockIdx.x * 32;
  auto fetch = [&](int buf, int k0) {
    for (int c = 0; c < 32; ++c) {
      const int bc = min(b_first + c, a.batch - 1);
      const int k = min(k0 + lane, a.n - 1);
      const double *src = a.carr + (size_t)bc * a.carr_stride + k;
      const unsigned dst = (unsigned)__cvta_generic_to_shared(&tile[buf][c][lane]);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  constexpr int HALF_PI_HI = 0x3ff921fb;
  constexpr unsigned HALF_PI_LO = 0x54442d18u;
  constexpr double ROUND_MAGIC = 6755399441055744.0;  // 1.5 * 2^52: x + M - M = rint(x) for |x| < 2^51
  double nd = n0;  // trigOffset, an exact integer
  fetch(0, 0);
  for (int k0 = 0, buf = 0; k0 < a.n; k0 += 32, buf ^= 1) {
    if (k0 + 32 < a.n) {
      fetch(buf ^ 1, k0 + 32);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    const double *xs = tile[buf][lane];
    const int kn = min(32, a.n - k0);
    // Only the sign of a sample enters the phase detector: reduce the tile to two bit masks
    // up front (independent loads and compares), so that the loop below touches nothing but
    // registers and no branch in it waits for a load.
    unsigned pos_mask = 0, zero_mask = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const double x = xs[j];
      pos_mask |= (x > 0.0 ? 1u : 0u) << j;
      zero_mask |= (x == 0.0 ? 1u : 0u) << j;
    }
    double *tp = th + k0;
    if (!__any_sync(0xffffffffu, zero_mask != 0)) {
#pragma unroll 8
      for (int j = 0; j < kn; ++j) {
        // off the chain: 2*pi*(f/Fs)*trigOffset (fmSupportLib.py:338-340)
        nd = __dadd_rn(nd, 1.0);
        const double wn = __dmul_rn(w, nd);
        // Phase detector (fmSupportLib.py:324-329), see the header.  The sign of r is read
        // from its high word: an integer test, not a compare on the double-precision pipe.
        const double c = (pos_mask >> j) & 1u ? 0.0 : (__double2hiint(r) < 0 ? -RDS_PI : RDS_PI);
        const double errorD = __dsub_rn(c, r);
        // loop filter and phase estimate (fmSupportLib.py:332-335).  The products are fused
        // into the sums (the model rounds them first: a difference of ~1e-21 per sample).
        integrator = fma(Ki, errorD, integrator);
        phaseEst = __dadd_rn(fma(Kp, errorD, phaseEst), integrator);
        const double trigArg = __dadd_rn(wn, phaseEst);
        tp[j] = trigArg;
        // reduce modulo 2*pi; the turn count is rounded with the magic constant (two 8-cycle
        // operations where cvt.rni.f64 takes 21, tools/ubench_dp_latency.cu)
        const double turns = __dsub_rn(fma(trigArg, RDS_INV_2PI, ROUND_MAGIC), ROUND_MAGIC);
        r = fma(-turns, RDS_2PI_LO, fma(-turns, RDS_2PI_HI, trigArg));
      }
    } else {
      // some capture of this warp has exact zeros in the tile (silence): same step with
      // atan2's signed-zero cases
      for (int j = 0; j < kn; ++j) {
        nd = __dadd_rn(nd, 1.0);
        const double wn = __dmul_rn(w, nd);
        const bool r_neg = __double2hiint(r) < 0;
        double c = (pos_mask >> j) & 1u ? 0.0 : (r_neg ? -RDS_PI : RDS_PI);
        double rr = r;
        if ((zero_mask >> j) & 1u) {
          // atan2(+-0, +-0): 0 when cos(t) has a clear sign bit (|t| <= pi/2), else +-pi by the
          // sign of -sin(t); |t| compared as an integer (exact for doubles)
          const int r_abs = __double2hiint(r) & 0x7fffffff;
          const bool far = r_abs > HALF_PI_HI || (r_abs == HALF_PI_HI && (unsigned)__double2loint(r) > HALF_PI_LO);
          c = far ? (r_neg ? RDS_PI : -RDS_PI) : 0.0;
          rr = 0.0;
        }
        const double errorD = __dsub_rn(c, rr);
        integrator = fma(Ki, errorD, integrator);
        phaseEst = __dadd_rn(fma(Kp, errorD, phaseEst), integrator);
        const double trigArg = __dadd_rn(wn, phaseEst);
        tp[j] = trigArg;
        const double turns = __dsub_rn(fma(trigArg, RDS_INV_2PI, ROUND_MAGIC), ROUND_MAGIC);
        r = fma(-turns, RDS_2PI_LO, fma(-turns, RDS_2PI_HI, trigArg));
      }
    }
    __syncwarp();
  }
  st[0] = integrator;
  st[1] = phaseEst;
  st[2] = r;
  st[5] = nd;
}

// ---------------------------------------------------------------------------
// R3a, warp form: one WARP per capture, lanes = 32 consecutive samples.
// With the phase detector written as e = c - (t mod 2*pi) (header of k_rds_pll) one step is
//     I' = I + Ki*e,   P' = P + Kp*e + I',   e = g - P,   g = c - r + P^
// where r and c are evaluated from a guess P^ of the previous sample's phase estimate: for a
// fixed set of discrete choices (sign of the sample, sign of the reduced phase, turn count,
// rounding of t) the step is AFFINE in (I, P) with a constant matrix
//     A = [[1, -Ki], [1, 1 - Kp - Ki]],   offset g * (Ki, Kp + Ki).
// So a tile of 32 samples is solved by a weighted prefix sum over the lanes (five shuffle
// rounds with A^1, A^2, .. A^16, then A^(j+1) times the tile's start state), the discrete
// choices are re-derived from the solved trajectory, and the tile is solved again until they
// stop changing -- every pass fixes at least the first sample whose choice was wrong, two
// passes are the rule, and a final solve with the converged choices makes the rounding of t
// consistent too.  Coalesced loads and stores.  (Described for one sample per lane; the kernel
// below puts two on a lane.)  Same model, same parity bound (1e-9 at every later stage).
// ---------------------------------------------------------------------------
struct RdsM2 {
  double a, b, c, d;  // [[a, b], [c, d]]
};
__device__ __forceinline__ RdsM2 rds_mm(const RdsM2 &x, const RdsM2 &y) {
  return {fma(x.a, y.a, x.b * y.c), fma(x.a, y.b, x.b * y.d), fma(x.c, y.a, x.d * y.c), fma(x.c, y.b, x.d * y.d)};
}

constexpr int RDS_PLLW_WARPS = 4;

// Two consecutive samples per lane, 64 per tile: a lane composes its two steps locally
// (A*b0 + b1 under A^2), the prefix sum runs over A^2, A^4, .. A^32, and the state between a
// lane's two samples is A times the previous lane's result plus b0.  Half the shuffle rounds
// per sample of a 32-sample tile.
static __global__ void __launch_bounds__(32 * RDS_PLLW_WARPS) k_rds_pll_warp(const RdsPllArgs a) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * RDS_PLLW_WARPS + (threadIdx.x >> 5);
  if (b >= a.batch) return;
  const unsigned FULL = 0xffffffffu;
  const double Kp = __dmul_rn(a.normBandwidth, 2.666);                                  // fmSupportLib.py:303-309
  const double Ki = __dmul_rn(__dmul_rn(a.normBandwidth, a.normBandwidth), 3.555);
  const double w = __dmul_rn(__dmul_rn(2.0, RDS_PI), __ddiv_rn(a.freq, a.Fs));          // fmSupportLib.py:340
  const double gI = Ki, gP = Kp + Ki;
  // pw[t] = A^(2^t); this lane's M = A^(2*(lane+1))
  RdsM2 pw[7];
  pw[0] = {1.0, -Ki, 1.0, 1.0 - Kp - Ki};
#pragma unroll
  for (int t = 1; t < 7; ++t) pw[t] = rds_mm(pw[t - 1], pw[t - 1]);
  const RdsM2 A = pw[0];
  RdsM2 M = {1.0, 0.0, 0.0, 1.0};
#pragma unroll
  for (int t = 1; t < 7; ++t)
    if (((2 * (lane + 1)) >> t) & 1) M = rds_mm(pw[t], M);
  double *st = a.state + (size_t)b * 8;
  double I0 = st[0], P0 = st[1], r0 = st[2];
  const double n0 = st[5];
  const double *in = a.carr + (size_t)b * a.carr_stride;
  double *th = a.theta + (size_t)b * a.theta_stride + 1;
  constexpr int HALF_PI_HI = 0x3ff921fb;
  constexpr unsigned HALF_PI_LO = 0x54442d18u;
  constexpr double ROUND_MAGIC = 6755399441055744.0;  // 1.5 * 2^52
  auto load2 = [&](int k, double &v0, double &v1) {
    v0 = k < a.n ? in[k] : 1.0;          // (past the end: any non-zero value, never used)
    v1 = k + 1 < a.n ? in[k + 1] : 1.0;
  };
  double xn0, xn1;
  load2(2 * lane, xn0, xn1);
  for (int k0 = 0; k0 < a.n; k0 += 64) {
    const int kn = min(64, a.n - k0);
    const double x0 = xn0, x1 = xn1;
    load2(k0 + 64 + 2 * lane, xn0, xn1);   // next tile, a tile ahead
    const bool v0 = 2 * lane < kn, v1 = 2 * lane + 1 < kn;
    // A tile with an exactly-zero sample (silence) is walked in order by every lane alike, with
    // the model's own sequence of operations: there the detector's output (0 or +-pi, atan2's
    // signed-zero cases) does not depend on the phase beyond its class, the phase estimate
    // grows without bound, and results become sensitive to the order of the roundings.
    if (__any_sync(FULL, (v0 && x0 == 0.0) || (v1 && x1 == 0.0))) {
      double I = I0, P = P0, r = r0, mine0 = 0.0, mine1 = 0.0;
      for (int j = 0; j < kn; ++j) {
        const double xj = __shfl_sync(FULL, (j & 1) ? x1 : x0, j >> 1);
        const bool r_neg = __double2hiint(r) < 0;
        double c = xj > 0.0 ? 0.0 : (r_neg ? -RDS_PI : RDS_PI);
        double rr = r;
        if (xj == 0.0) {
          const int r_abs = __double2hiint(r) & 0x7fffffff;
          const bool far = r_abs > HALF_PI_HI || (r_abs == HALF_PI_HI && (unsigned)__double2loint(r) > HALF_PI_LO);
          c = far ? (r_neg ? RDS_PI : -RDS_PI) : 0.0;
          rr = 0.0;
        }
        const double errorD = __dsub_rn(c, rr);
        I = fma(Ki, errorD, I);
        P = __dadd_rn(fma(Kp, errorD, P), I);
        const double t = __dadd_rn(__dmul_rn(w, n0 + (double)(k0 + j + 1)), P);
        if (lane == (j >> 1)) {
          if (j & 1) mine1 = t;
          else mine0 = t;
        }
        const double turns = __dsub_rn(fma(t, RDS_INV_2PI, ROUND_MAGIC), ROUND_MAGIC);
        r = fma(-turns, RDS_2PI_LO, fma(-turns, RDS_2PI_HI, t));
      }
      if (v0) th[k0 + 2 * lane] = mine0;
      if (v1) th[k0 + 2 * lane + 1] = mine1;
      I0 = I;
      P0 = P;
      r0 = r;
      continue;
    }
    // this lane's samples are k = k0 + 2*lane and k + 1; a step uses t of the sample before it
    const double wn_a = __dmul_rn(w, n0 + (double)(k0 + 2 * lane));       // t of sample k-1
    const double wn_b = __dmul_rn(w, n0 + (double)(k0 + 2 * lane + 1));   // t of sample k
    const double wn_c = __dmul_rn(w, n0 + (double)(k0 + 2 * lane + 2));   // t of sample k+1
    double Pp0 = lane == 0 ? P0 : fma((double)(2 * lane), I0, P0);        // first guess: the integrator's drift
    double Pp1 = fma((double)(2 * lane + 1), I0, P0);
    double SI0 = 0.0, SP0 = 0.0, SI1 = 0.0, SP1 = 0.0;                    // state after sample k / k+1
    double sg_t0 = 0.0, sg_t1 = 0.0;
    int sg_b0 = -1, sg_b1 = -1;
    for (int it = 0; it < 72; ++it) {
      // ---- discrete choices and g = c - r + P^ from the guesses ----
      double ra, tu0 = 0.0;
      if (lane == 0) {
        ra = r0;  // exact: carried from the previous tile / call
      } else {
        const double t = __dadd_rn(wn_a, Pp0);
        tu0 = __dsub_rn(fma(t, RDS_INV_2PI, ROUND_MAGIC), ROUND_MAGIC);
        ra = fma(-tu0, RDS_2PI_LO, fma(-tu0, RDS_2PI_HI, t));
      }
      const double tb = __dadd_rn(wn_b, Pp1);
      const double tu1 = __dsub_rn(fma(tb, RDS_INV_2PI, ROUND_MAGIC), ROUND_MAGIC);
      const double rb = fma(-tu1, RDS_2PI_LO, fma(-tu1, RDS_2PI_HI, tb));
      const int n0b = __double2hiint(ra) < 0, n1b = __double2hiint(rb) < 0;
      const double c0 = x0 > 0.0 ? 0.0 : (n0b ? -RDS_PI : RDS_PI);
      const double c1 = x1 > 0.0 ? 0.0 : (n1b ? -RDS_PI : RDS_PI);
      const bool same = it > 0 && (!v0 || (n0b == sg_b0 && tu0 == sg_t0)) && (!v1 || (n1b == sg_b1 && tu1 == sg_t1));
      const bool converged = __all_sync(FULL, same);
      sg_b0 = n0b; sg_b1 = n1b; sg_t0 = tu0; sg_t1 = tu1;
      const double g0 = v0 ? __dadd_rn(__dsub_rn(c0, ra), Pp0) : 0.0;
      const double g1 = v1 ? __dadd_rn(__dsub_rn(c1, rb), Pp1) : 0.0;
      // ---- solve: per lane (I, P) -> A^2 (I, P) + (A b0 + b1), b = g (Ki, Kp + Ki) ----
      const double b0I = g0 * gI, b0P = g0 * gP;
      double vI = fma(A.a, b0I, fma(A.b, b0P, g1 * gI));
      double vP = fma(A.c, b0I, fma(A.d, b0P, g1 * gP));
#pragma unroll
      for (int t = 0; t < 5; ++t) {
        const double uI = __shfl_up_sync(FULL, vI, 1 << t), uP = __shfl_up_sync(FULL, vP, 1 << t);
        if (lane >= (1 << t)) {
          vI = fma(pw[t + 1].a, uI, fma(pw[t + 1].b, uP, vI));
          vP = fma(pw[t + 1].c, uI, fma(pw[t + 1].d, uP, vP));
        }
      }
      SI1 = fma(M.a, I0, fma(M.b, P0, vI));
      SP1 = fma(M.c, I0, fma(M.d, P0, vP));
      double pI = __shfl_up_sync(FULL, SI1, 1), pP = __shfl_up_sync(FULL, SP1, 1);
      if (lane == 0) { pI = I0; pP = P0; }
      SI0 = fma(A.a, pI, fma(A.b, pP, b0I));
      SP0 = fma(A.c, pI, fma(A.d, pP, b0P));
      Pp0 = pP;    // P of the sample before this lane's first
      Pp1 = SP0;   // P of this lane's first sample
      if (converged) break;  // this pass was solved with settled choices and a settled guess
    }
    // ---- this tile's NCO phases; state for the next tile ----
    const double t0 = __dadd_rn(wn_b, SP0), t1 = __dadd_rn(wn_c, SP1);
    if (v0) th[k0 + 2 * lane] = t0;
    if (v1) th[k0 + 2 * lane + 1] = t1;
    const int last = kn - 1, ll = last >> 1;
    const bool odd = last & 1;
    I0 = __shfl_sync(FULL, odd ? SI1 : SI0, ll);
    P0 = __shfl_sync(FULL, odd ? SP1 : SP0, ll);
    const double tl = __shfl_sync(FULL, odd ? t1 : t0, ll);
    const double turns_l = __dsub_rn(fma(tl, RDS_INV_2PI, ROUND_MAGIC), ROUND_MAGIC);
    r0 = fma(-turns_l, RDS_2PI_LO, fma(-turns_l, RDS_2PI_HI, tl));
  }
  if (lane == 0) {
    st[0] = I0;
    st[1] = P0;
    st[2] = r0;
    st[5] = n0 + (double)a.n;
  }
}

// ---------------------------------------------------------------------------
// R3b: NCO outputs cos/sin(t*ncoScale + phaseAdjust) (fmSupportLib.py:344-345), the all-pass
// delay of the channel signal (fmRDS.py:227) and both mixers (fmRDS.py:241,251), one thread
// per sample.  ncoOut[0] of a call is the last NCO value of the previous call; the mixers use
// ncoOut[0..N), i.e. the NCO delayed by one sample, so sample k uses theta[k] (theta[0] being
// the carried phase).  Before the first sample both NCO outputs are 1.0 (fmRDS.py:175).
// ---------------------------------------------------------------------------
struct RdsMixArgs {
  const double *theta;
  size_t theta_stride;
  const double *chan;  // [B][chan_stride], sample 0 at chan_off
  size_t chan_stride;
  int chan_off;
  double *mixI, *mixQ;  // [B][mix_stride], sample 0 at mix_off
  size_t mix_stride;
  int mix_off;
  double *ncoI, *ncoQ;  // optional [B][nco_stride]: ncoOut[0..N] for the parity taps
  size_t nco_stride;
  int n;
  double ncoScale, phaseAdjust;
};

static __global__ void __launch_bounds__(256) k_rds_mix(const RdsMixArgs a) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (k > a.n || (k == a.n && !a.ncoI)) return;
  const double t = a.theta[(size_t)b * a.theta_stride + k];
  double oI = 1.0, oQ = 1.0;
  if (t == t) {
    const double arg = __dadd_rn(__dmul_rn(t, a.ncoScale), a.phaseAdjust);
    sincos(rds_reduce_2pi(arg), &oQ, &oI);
  }
  if (a.ncoI) {
    a.ncoI[(size_t)b * a.nco_stride + k] = oI;
    a.ncoQ[(size_t)b * a.nco_stride + k] = oQ;
  }
  if (k < a.n) {
    const double d = a.chan[(size_t)b * a.chan_stride + a.chan_off - RDS_DELAY + k];
    a.mixI[(size_t)b * a.mix_stride + a.mix_off + k] = __dmul_rn(__dmul_rn(oI, d), 2.0);
    a.mixQ[(size_t)b * a.mix_stride + a.mix_off + k] = __dmul_rn(__dmul_rn(oQ, d), 2.0);
  }
}

// ---------------------------------------------------------------------------
// R4: rational resampler (fmSupportLib.py:388-407): output j takes phase (jD mod U) of the
// 101*U-tap low-pass and the 101 inputs ending at floor(jD/U); gain U.
// I and Q share the taps.
// ---------------------------------------------------------------------------
struct RdsResampleArgs {
  const double *mixI, *mixQ;
  size_t mix_stride;
  int mix_off;
  double *rsI, *rsQ;
  size_t rs_stride;
  int rs_off;
  const double *quad;  // [U][quad_rows][4]
  int U, D, n_out;
};

// Lanes are captures and a warp computes FOUR consecutive outputs at a time for 32 captures.
// The four outputs' input windows overlap (consecutive outputs start D/U = 3.9 or 2.4 samples
// apart and each reaches 101 samples back), so the warp walks the union of the windows once,
// newest sample first: one 64-bit shared-memory read per sample and component feeds four
// accumulators.  The four taps that meet a sample come from a host-built table indexed by the
// phase of the quad's first output, [U][rows][4], laid out in walk order with zeros where a
// sample lies outside an output's window -- warp-uniform 16-byte loads.  Every accumulator
// still sees its 101 products in ascending tap order.  (The first version, one output per
// warp pass, spent 69 % of the shared-memory/L1 data pipe on 10 % of the FP64 pipe: 2.27 ms.)
// The input tile is stored [time][capture] (pitch 33: transposing fill and per-capture reads
// are both conflict-free) and filled with asynchronous 8-byte copies; a block owns 16
// consecutive outputs of 32 captures and writes its results back through shared memory so
// that global stores are contiguous.
constexpr int RDS_RS_J = 16;      // outputs per block (4 quads, one per warp)
constexpr int RDS_RS_PITCH = 33;  // doubles per tile row

static __global__ void __launch_bounds__(256) k_rds_resample(const RdsResampleArgs a, int batch, int rows_cap,
                                                             int n_in, int quad_rows) {
  extern __shared__ __align__(16) double rs_sm[];
  double *tI = rs_sm;
  double *tQ = tI + (size_t)rows_cap * RDS_RS_PITCH;
  double *oI = tQ + (size_t)rows_cap * RDS_RS_PITCH;  // [32][J+1]
  double *oQ = oI + 32 * (RDS_RS_J + 1);
  // [4 quads][quad_rows][2], on the next 16-byte boundary
  double2 *tabs = reinterpret_cast<double2 *>(rs_sm + (((size_t)2 * rows_cap * RDS_RS_PITCH + 2 * 32 * (RDS_RS_J + 1) + 1) & ~(size_t)1));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j0 = blockIdx.x * RDS_RS_J;
  // this block's four table entries, fetched alongside the tile (read in place they cost one
  // L2 round trip per 128-byte line inside the tap loop: 3.5 ms instead of 2.3 ms)
  for (int i = threadIdx.x; i < 4 * quad_rows * 2; i += 256) {
    const int w = i / (quad_rows * 2), e = i % (quad_rows * 2);
    const int p0w = (int)(((long long)(j0 + 4 * w) * a.D) % a.U);
    const double2 *src = reinterpret_cast<const double2 *>(a.quad) + (size_t)p0w * quad_rows * 2 + e;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(tabs + i)), "l"(src) : "memory");
  }
  const int b0 = blockIdx.y * 32;
  const long long base0 = ((long long)j0 * a.D) / a.U;                        // newest input of the first output
  const long long baseL = ((long long)(j0 + RDS_RS_J - 1) * a.D) / a.U;       // ... of the last one (may lie past the data)
  const long long lo = base0 - (RDS_TP - 1);                                  // oldest input needed
  const int rows = (int)(baseL - lo + 1);
  for (int c = warp; c < 32; c += 8) {
    const int bc = min(b0 + c, batch - 1);
    const double *sI = a.mixI + (size_t)bc * a.mix_stride + a.mix_off + lo;
    const double *sQ = a.mixQ + (size_t)bc * a.mix_stride + a.mix_off + lo;
    for (int q = lane; q < rows; q += 32) {
      double *dI = tI + q * RDS_RS_PITCH + c, *dQ = tQ + q * RDS_RS_PITCH + c;
      if (lo + q < n_in) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dI)), "l"(sI + q) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dQ)), "l"(sQ + q) : "memory");
      } else {  // past the end of the call: only zero taps of outputs that are not stored meet these
        *dI = 0.0;
        *dQ = 0.0;
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // warps 0-3 take the I component of quads 0-3, warps 4-7 the Q component: twice the warps per
  // tile to hide the shared-memory latency of the tap loop
  const int quad = warp & 3;
  const bool is_q = warp >= 4;
  const int jq = j0 + 4 * quad;  // first output of this warp's quad
  if (jq < a.n_out) {
    const long long m = (long long)jq * a.D;
    const int p0 = (int)(m % a.U);
    const int span = (p0 + 3 * a.D) / a.U;                 // newest input of output 3 - that of output 0
    const int top = (int)(m / a.U + span - lo);            // tile row of the newest input of output 3
    const int n_rows = RDS_TP + span;
    const double2 *tq = tabs + (size_t)quad * quad_rows * 2;
    const double *x = (is_q ? tQ : tI) + (size_t)top * RDS_RS_PITCH + lane;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 8
    for (int r = 0; r < n_rows; ++r) {
      const double2 h01 = tq[2 * r], h23 = tq[2 * r + 1];
      const double v = x[-r * RDS_RS_PITCH];
      acc[0] = fma(h01.x, v, acc[0]);
      acc[1] = fma(h01.y, v, acc[1]);
      acc[2] = fma(h23.x, v, acc[2]);
      acc[3] = fma(h23.y, v, acc[3]);
    }
    double *o = is_q ? oQ : oI;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      o[lane * (RDS_RS_J + 1) + 4 * quad + i] = __dmul_rn(acc[i], (double)a.U);  // fmSupportLib.py:400
  }
  __syncthreads();
  const int jn = min(RDS_RS_J, a.n_out - j0);
  for (int i = threadIdx.x; i < 32 * RDS_RS_J; i += 256) {
    const int c = i / RDS_RS_J, jj = i % RDS_RS_J;
    if (b0 + c < batch && jj < jn) {
      a.rsI[(size_t)(b0 + c) * a.rs_stride + a.rs_off + j0 + jj] = oI[c * (RDS_RS_J + 1) + jj];
      a.rsQ[(size_t)(b0 + c) * a.rs_stride + a.rs_off + j0 + jj] = oQ[c * (RDS_RS_J + 1) + jj];
    }
  }
}

// ---------------------------------------------------------------------------
// R6: clock and data recovery + Manchester decoding (fmSupportLib.py:103-222), driven as in
// fmRDS.py:257-268: per block, pair = (0,0), start = 158, prev_size = 0 (which makes the
// model's "pair with the previous block" branch, :117-125, dead).  One warp per
// (capture, block).  The model builds the whole array of sampling points, then walks the
// pairs; a pair of equal signs that cannot be repaired by inverting a small sample moves
// the start by one symbol and starts over.  Pair decisions only depend on the two samples of
// the pair, so one pass can decide and emit bits while it samples; a restart rewinds the
// output to the bits the restarts themselves produced.
// Where the model would never leave its loop (a whole pass without a single pair of opposite
// signs, e.g. fewer than two sampling points left), the pass is accepted as it stands.
// ---------------------------------------------------------------------------
struct RdsCdrArgs {
  const double *rrc;  // [B][rrc_stride]
  size_t rrc_stride;
  int block_out;      // samples per block
  int n_blocks;       // blocks in this call
  int sps;
  long long first_block;  // index of this call's first block since reset (block_count, :157)
  uint8_t *bits;      // [B][blocks_cap][bits_cap]
  int *counts;        // [B][blocks_cap]
  int blocks_cap, bits_cap, cursor;
  int batch;
};

// A restart moves the start by exactly one symbol, so every pass looks at the same grid of
// sampling points x[158 + i*sps], minus its first few: the warp fetches the grid once (lanes
// in parallel) into shared memory and lane 0 walks it.
constexpr int RDS_CDR_WARPS = 4;
constexpr int RDS_CDR_PTS = 512;  // sampling points per block kept in shared memory (else read in place)

// The walk over one block's grid of sampling points pts[i * step], i < n_pts (lane 0 only).
// pair0 is the model's pair[0]; on return `size` is the number of points of the accepted pass,
// `first_pt` the index of its first point and `last` the model's samples[-1] (after repairs).
struct RdsCdrWalk {
  int nb, size, first_pt;
  double pair0, last;
};

static __device__ RdsCdrWalk rds_cdr_walk(const double *pts, int step, int n_pts, bool first_ever, double pair0,
                                          int n_prefix, uint8_t *out, int bits_cap) {
  const double limit = 0.3;
  int first_pt = 0;  // restarts so far: the pass starts `first_pt` symbols later
  int nb = 0;
  double last = 0.0;
  for (;;) {
    double p1 = 0.0, p2 = 0.0;  // the two previous sampling points (before pair repairs)
    double first = 0.0, s0 = 0.0;
    bool restart = false;
    nb = n_prefix;
    for (int k = 0; first_pt + k < n_pts; ++k) {
      const double xi = pts[(size_t)(first_pt + k) * step];
      double s = xi;
      // :128-136 a third consecutive high (or low) is inverted
      if (k >= 2 && ((p2 > 0 && p1 > 0 && xi > 0) || (p2 < 0 && p1 < 0 && xi < 0))) s = -xi;
      p2 = p1;
      p1 = s;
      if ((k & 1) == 0) {
        first = s;
        last = s;
        if (k == 0) s0 = s;
        continue;
      }
      double u = first, v = s;
      if ((u < 0 && v < 0) || (u > 0 && v > 0)) {  // :147
        if (fabs(u) < limit) u = -u;                // :151-153
        else if (fabs(v) < limit) v = -v;           // :154-156
        else {                                      // :158-172
          restart = true;
          break;
        }
      }
      if (k == 1) s0 = u;  // samples[0] as the model sees it at a later restart
      last = v;
      // manchestering, :203-222
      uint8_t bit = 0;
      if (u > 0 && v < 0) bit = 1;
      if (nb < bits_cap) out[nb] = bit;
      ++nb;
    }
    if (!restart) break;
    first_pt += 1;
    if (!first_ever) {  // :160-167: symbolToBit looks at pair[0] only (:228-236)
      const uint8_t bit = pair0 > 0 ? 1 : 0;
      if (n_prefix < bits_cap) out[n_prefix] = bit;
      ++n_prefix;
      pair0 = s0;
    }
  }
  RdsCdrWalk w;
  w.nb = nb;
  w.size = max(n_pts - first_pt, 0);
  w.first_pt = first_pt;
  w.pair0 = pair0;
  w.last = last;
  return w;
}

static __global__ void __launch_bounds__(32 * RDS_CDR_WARPS) k_rds_cdr(const RdsCdrArgs a) {
  __shared__ double pts_sm[RDS_CDR_WARPS][RDS_CDR_PTS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int idx = blockIdx.x * RDS_CDR_WARPS + warp;
  if (idx >= a.batch * a.n_blocks) return;
  const int b = idx / a.n_blocks, blk = idx % a.n_blocks;
  const double *x = a.rrc + (size_t)b * a.rrc_stride + (size_t)blk * a.block_out + RDS_CDR_START;
  const int sps = a.sps;
  const int n_pts = (a.block_out - RDS_CDR_START + sps - 1) / sps;  // points i with 158 + i*sps < n
  const double *pts = x;
  int step = sps;
  if (n_pts <= RDS_CDR_PTS) {
    for (int i = lane; i < n_pts; i += 32) pts_sm[warp][i] = x[(size_t)i * sps];
    __syncwarp();
    pts = pts_sm[warp];
    step = 1;
  }
  if (lane != 0) return;
  uint8_t *out = a.bits + ((size_t)b * a.blocks_cap + a.cursor + blk) * a.bits_cap;
  const RdsCdrWalk w = rds_cdr_walk(pts, step, n_pts, (a.first_block + blk) == 0, 0.0, 0, out, a.bits_cap);
  a.counts[(size_t)b * a.blocks_cap + a.cursor + blk] = min(w.nb, a.bits_cap);
}

// The same with the model's to_pass_on_state {pair[0], start, prev_size} carried from block to
// block (fmSupportLib.py:104-106,178-189) instead of re-created per block: one warp per capture
// walks its blocks in order.  An odd number of points in the previous block leaves one symbol
// that is paired with the first point of this block (:117-125).  cdr_state: [B][4] doubles =
// pair[0], start, prev_size, unused.
static __global__ void __launch_bounds__(32 * RDS_CDR_WARPS) k_rds_cdr_carry(const RdsCdrArgs a, double *cdr_state) {
  __shared__ double pts_sm[RDS_CDR_WARPS][RDS_CDR_PTS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * RDS_CDR_WARPS + warp;
  if (b >= a.batch) return;
  double *st = cdr_state + (size_t)b * 4;
  double pair0 = st[0];
  int start = (int)st[1], prev_size = (int)st[2];
  const int sps = a.sps, n = a.block_out;
  for (int blk = 0; blk < a.n_blocks; ++blk) {
    const double *x = a.rrc + (size_t)b * a.rrc_stride + (size_t)blk * n;
    uint8_t *out = a.bits + ((size_t)b * a.blocks_cap + a.cursor + blk) * a.bits_cap;
    int n_prefix = 0;
    if ((prev_size & 1) && start < n) {  // :117-125
      if (lane == 0 && a.bits_cap > 0) out[0] = pair0 > 0 ? 1 : 0;
      n_prefix = 1;
      pair0 = x[start];
      start += sps;
    }
    const int n_pts = start < n ? (n - start + sps - 1) / sps : 0;
    const double *pts = x + start;
    int step = sps;
    __syncwarp();
    if (n_pts <= RDS_CDR_PTS) {
      for (int i = lane; i < n_pts; i += 32) pts_sm[warp][i] = pts[(size_t)i * sps];
      __syncwarp();
      pts = pts_sm[warp];
      step = 1;
    }
    RdsCdrWalk w{};
    if (lane == 0) {
      w = rds_cdr_walk(pts, step, n_pts, (a.first_block + blk) == 0, pair0, n_prefix, out, a.bits_cap);
      a.counts[(size_t)b * a.blocks_cap + a.cursor + blk] = min(w.nb, a.bits_cap);
    }
    // :178-189 state for the next block, computed by lane 0 and shared with the warp
    const int size = __shfl_sync(0xffffffffu, w.size, 0);
    const int first_pt = __shfl_sync(0xffffffffu, w.first_pt, 0);
    const double p0 = __shfl_sync(0xffffffffu, w.pair0, 0), last = __shfl_sync(0xffffffffu, w.last, 0);
    pair0 = size > 0 ? last : p0;
    const int last_index = (size - 1) * sps + (start + first_pt * sps);
    start = sps - (n - last_index);
    prev_size = size;
  }
  if (lane == 0) {
    st[0] = pair0;
    st[1] = (double)start;
    st[2] = (double)prev_size;
  }
}

// ---------------------------------------------------------------------------
// R7: move the tails into the history prefixes; keep the last demod samples.
// ---------------------------------------------------------------------------
struct RdsCarryArgs {
  double *rows[6];
  size_t strides[6];
  int src_off[6];  // first element to copy
  int len[6];
  const float *demod;  // this call's fm_demod, sample 0 at demod_off
  size_t demod_stride;
  int demod_off;
  int n_if;
  float *hist32;  // [B][hist_len]
  int hist_len;
};

static __global__ void k_rds_carry(const RdsCarryArgs c) {
  __shared__ double stage[RDS_HC + 8];
  const int b = blockIdx.x, t = threadIdx.x;
  for (int r = 0; r < 6; ++r) {
    if (!c.rows[r]) continue;
    double *row = c.rows[r] + (size_t)b * c.strides[r];
    for (int i = t; i < c.len[r]; i += blockDim.x) stage[i] = row[c.src_off[r] + i];
    __syncthreads();
    for (int i = t; i < c.len[r]; i += blockDim.x) row[i] = stage[i];
    __syncthreads();
  }
  // last hist_len samples of (old history ++ this call's fm_demod)
  float *fs = reinterpret_cast<float *>(stage);
  float *h = c.hist32 + (size_t)b * c.hist_len;
  const float *d = c.demod + (size_t)b * c.demod_stride + c.demod_off;
  for (int i = t; i < c.hist_len; i += blockDim.x) {
    const int j = c.n_if - c.hist_len + i;
    fs[i] = j >= 0 ? d[j] : h[c.hist_len + j];
  }
  __syncthreads();
  for (int i = t; i < c.hist_len; i += blockDim.x) h[i] = fs[i];
}

// ---------------------------------------------------------------------------
// host: filter design in double, like the model
// ---------------------------------------------------------------------------
// fmSupportLib.py:358-372
static std::vector<double> rds_band_pass(int ntaps, double Fs, double Fb, double Fe) {
  std::vector<double> h((size_t)ntaps);
  const double center = ((Fe + Fb) / 2) / (Fs / 2);
  const double pass = (Fe - Fb) / (Fs / 2);
  const double mid = (double)(ntaps - 1) / 2;
  for (int i = 0; i < ntaps; ++i) {
    double v = pass;
    if ((double)i != mid) {
      const double arg = M_PI * pass / 2 * ((double)i - mid);
      v = pass * (std::sin(arg) / arg);
    }
    v = v * std::cos((double)i * M_PI * center);
    const double win = std::sin((double)i * M_PI / (double)ntaps);
    h[(size_t)i] = v * (win * win);
  }
  return h;
}

// fmSupportLib.py:376-385
static std::vector<double> rds_low_pass(int ntaps, double Fs, double Fc) {
  std::vector<double> h((size_t)ntaps);
  const double norm = Fc / (Fs / 2);
  const double mid = (double)(ntaps - 1) / 2;
  for (int i = 0; i < ntaps; ++i) {
    double v = norm;
    if ((double)i != mid) {
      const double arg = M_PI * norm * ((double)i - mid);
      v = norm * (std::sin(arg) / arg);
    }
    const double win = std::sin((double)i * M_PI / (double)ntaps);
    h[(size_t)i] = v * (win * win);
  }
  return h;
}

// fmSupportLib.py:251-287
static std::vector<double> rds_rrc(double Fs, int ntaps) {
  std::vector<double> h((size_t)ntaps);
  const double Ts = 1 / 2375.0, beta = 0.90;
  for (int k = 0; k < ntaps; ++k) {
    const double t = ((double)k - (double)ntaps / 2) / Fs;
    double v;
    if (t == 0.0) {
      v = 1.0 + beta * ((4 / M_PI) - 1);
    } else if (t == -Ts / (4 * beta) || t == Ts / (4 * beta)) {
      v = (beta / std::sqrt(2.0)) * (((1 + 2 / M_PI) * std::sin(M_PI / (4 * beta))) +
                                     ((1 - 2 / M_PI) * std::cos(M_PI / (4 * beta))));
    } else {
      const double q = 4 * beta * t / Ts;
      v = (std::sin(M_PI * t * (1 - beta) / Ts) + 4 * beta * (t / Ts) * std::cos(M_PI * t * (1 + beta) / Ts)) /
          (M_PI * t * (1 - q * q) / Ts);
    }
    h[(size_t)k] = v;
  }
  return h;
}

template <typename T>
struct RBuf {
  T *p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    release();
    n = count;
    if (!count) return SDR_OK;
    SDR_CUDA(cudaMalloc(&p, count * sizeof(T)));
    SDR_CUDA(cudaMemset(p, 0, count * sizeof(T)));
    return SDR_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~RBuf() { release(); }
};

}  // namespace sdr

struct sdr_rds {
  sdr_pipeline *pipe = nullptr;
  DemodView view{};
  int mode = 0, U = 0, D = 0, sps = 0;
  int block_if = 0, block_out = 0, block_bytes = 0;
  int blocks_cap = 0, bits_cap = 0;
  int quad_rows = 0;  // rows of one entry of the resampler's quad table
  size_t cap_if = 0, cap_out = 0;
  size_t chan_stride = 0, carr_stride = 0, theta_stride = 0, mix_stride = 0, rs_stride = 0, rrc_stride = 0, nco_stride = 0;
  DTaps<RDS_T> h_chan{}, h_carr{};
  DTaps<RDS_TP> h_rrc{};
  DTaps<RDS_T, float> h_chan32{}, h_carr32{};   // the same coefficients rounded to float (precision = SDR_RDS_F32_FIR)
  DTaps<RDS_TP, float> h_rrc32{};
  int precision = 0;
  RBuf<double> cdr_state, quad, chan, carr, theta, mixI, mixQ, rsI, rsQ, rrcI, rrcQ, ncoI, ncoQ, pll;
  RBuf<float> hist32;
  RBuf<uint8_t> bits;
  RBuf<int> counts;
  bool keep_nco = false;
  bool cdr_carry = false;
  int pll_form = 0;   // SDR_RDS_PLL_AUTO / _LANE / _WARP
  int cursor = 0;             // blocks waiting in bits/counts
  long long blocks_done = 0;  // since reset
  size_t last_n_if = 0;
  cudaStream_t last_stream = nullptr;
  // host mirror of the pending results + the frame synchroniser's carried bits per capture
  bool mirrored = true;
  std::vector<uint8_t> h_bits;
  std::vector<int> h_counts;
  std::vector<std::vector<uint8_t>> decoded;  // fmRDS.py:272-276 decoded_data
  std::vector<std::string> offsets;           // per capture, one letter per pending block
  std::vector<std::vector<uint8_t>> diff;     // per capture, differential bits of pending blocks
};

static int rds_reset_device(sdr_rds *r) {
  SDR_CUDA(cudaSetDevice(r->view.device));
  const size_t B = (size_t)r->view.batch;
  double *rows[] = {r->chan.p, r->carr.p, r->mixI.p, r->mixQ.p, r->rsI.p, r->rsQ.p, r->rrcI.p, r->rrcQ.p};
  size_t sizes[] = {r->chan.n, r->carr.n, r->mixI.n, r->mixQ.n, r->rsI.n, r->rsQ.n, r->rrcI.n, r->rrcQ.n};
  for (int i = 0; i < 8; ++i)
    if (rows[i]) SDR_CUDA(cudaMemset(rows[i], 0, sizes[i] * sizeof(double)));
  SDR_CUDA(cudaMemset(r->hist32.p, 0, r->hist32.n * sizeof(float)));
  // fmRDS.py:175 state_rds_pll = [0, 0, 1, 0, 1, 0, 1]: integrator, phaseEst and trigOffset 0,
  // feedback phasor 1+0j (NCO phase 0), both NCO outputs 1.0 (theta[0] = NaN stands for that)
  SDR_CUDA(cudaMemset(r->pll.p, 0, r->pll.n * sizeof(double)));
  const double nan = std::nan("");
  std::vector<double> nans(B, nan);
  SDR_CUDA(cudaMemcpy2D(r->theta.p, r->theta_stride * sizeof(double), nans.data(), sizeof(double),
                        sizeof(double), B, cudaMemcpyHostToDevice));
  {  // fmRDS.py:257-260: pair = (0, 0), start = 158, prev_size = 0
    std::vector<double> cs(B * 4, 0.0);
    for (size_t b = 0; b < B; ++b) cs[b * 4 + 1] = (double)RDS_CDR_START;
    SDR_CUDA(cudaMemcpy(r->cdr_state.p, cs.data(), cs.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  r->cursor = 0;
  r->blocks_done = 0;
  r->last_n_if = 0;
  r->mirrored = true;
  for (auto &d : r->decoded) d.clear();
  for (auto &o : r->offsets) o.clear();
  for (auto &d : r->diff) d.clear();
  std::fill(r->h_counts.begin(), r->h_counts.end(), 0);
  return SDR_OK;
}

static int segs_for(int n, int tile, int batch, int z, int *outs_per_seg) {
  const int n_tiles = (n + tile - 1) / tile;
  // enough blocks for ~10 waves of 148 SMs x 12 resident CTAs: a grid of one or two waves
  // (the first version: 2048 blocks = 1.15 waves) leaves most SMs idle during the last one
  int want = (16384 + batch * z - 1) / (batch * z);
  want = std::max(1, std::min(want, n_tiles));
  const int per = (n_tiles + want - 1) / want;
  *outs_per_seg = per * tile;
  return (n_tiles + per - 1) / per;
}

static int rds_validate(const sdr_rds *r, size_t n_if) {
  if (n_if % (size_t)r->block_if) return fail(SDR_ERR_INVALID, "RDS: the call is not a whole number of RDS blocks");
  if (n_if > r->cap_if) return fail(SDR_ERR_CAPACITY, "RDS: call exceeds the capacity fixed at create time");
  if (r->cursor + (int)(n_if / (size_t)r->block_if) > r->blocks_cap)
    return fail(SDR_ERR_CAPACITY, "RDS: result buffer full; call sdr_rds_read, then sdr_rds_discard");
  return SDR_OK;
}

static int rds_process(sdr_rds *r, size_t n_if, cudaStream_t s) {
  if (n_if == 0) return SDR_OK;
  int rc = rds_validate(r, n_if);
  if (rc) return rc;
  const int n_blocks = (int)(n_if / (size_t)r->block_if);
  const int B = r->view.batch;
  const int n = (int)n_if;
  const int n_out = (int)((long long)n_if * r->U / r->D);
  sdr_pipeline *p = r->pipe;
  int ops;
  constexpr int R = 8, NT = 128, TILE = R * NT;

  {  // R1
    RdsFirArgs a{};
    a.src[0] = r->view.demod;
    a.src_stride = r->view.stride;
    a.src_off = r->view.off;
    a.hist32 = r->hist32.p;
    a.hist_len = RDS_T - 1;
    a.dst[0] = r->chan.p;
    a.dst_stride = r->chan_stride;
    a.dst_off = RDS_HC;
    a.n = n;
    dim3 grid(segs_for(n, TILE, B, 1, &ops), B, 1);
    a.outs_per_seg = ops;
    sdr_prof_begin(p, "k_rds_fir_channel", s);
    if (r->precision) k_rds_fir<RDS_T, R, NT, 0, float><<<grid, NT, 0, s>>>(a, r->h_chan32);
    else k_rds_fir<RDS_T, R, NT, 0><<<grid, NT, 0, s>>>(a, r->h_chan);
    if ((rc = sdr_check_launch(p, "k_rds_fir_channel"))) return rc;
  }
  {  // R2
    RdsFirArgs a{};
    a.src[0] = r->chan.p;
    a.src_stride = r->chan_stride;
    a.src_off = RDS_HC;
    a.dst[0] = r->carr.p;
    a.dst_stride = r->carr_stride;
    a.dst_off = 0;
    a.n = n;
    dim3 grid(segs_for(n, TILE, B, 1, &ops), B, 1);
    a.outs_per_seg = ops;
    sdr_prof_begin(p, "k_rds_fir_carrier", s);
    if (r->precision) k_rds_fir<RDS_T, R, NT, 1, float><<<grid, NT, 0, s>>>(a, r->h_carr32);
    else k_rds_fir<RDS_T, R, NT, 1><<<grid, NT, 0, s>>>(a, r->h_carr);
    if ((rc = sdr_check_launch(p, "k_rds_fir_carrier"))) return rc;
  }
  {  // R3a
    RdsPllArgs a{};
    a.carr = r->carr.p;
    a.carr_stride = r->carr_stride;
    a.theta = r->theta.p;
    a.theta_stride = r->theta_stride;
    a.state = r->pll.p;
    a.n = n;
    a.batch = B;
    // fmRDS.py:236-237
    a.freq = 114e3;
    a.Fs = (double)r->view.if_Fs;
    a.normBandwidth = 0.002;
    // Warp form below ~4096 captures (latency: 1.8 ms per 76 800 samples, flat up to ~1024
    // captures, then ~0.95 ms per 1024 captures); one lane per capture above (4.7 ms, flat up to
    // ~19 k captures): measured 8192 captures 7.9 vs 4.8 ms.  sdr_rds_config.pll_form forces one.
    const bool one_lane = r->pll_form ? r->pll_form == SDR_RDS_PLL_LANE : B > 4096;
    sdr_prof_begin(p, "k_rds_pll", s);
    if (one_lane) k_rds_pll<<<(B + 31) / 32, 32, 0, s>>>(a);
    else k_rds_pll_warp<<<(B + RDS_PLLW_WARPS - 1) / RDS_PLLW_WARPS, 32 * RDS_PLLW_WARPS, 0, s>>>(a);
    if ((rc = sdr_check_launch(p, "k_rds_pll"))) return rc;
  }
  {  // R3b
    RdsMixArgs a{};
    a.theta = r->theta.p;
    a.theta_stride = r->theta_stride;
    a.chan = r->chan.p;
    a.chan_stride = r->chan_stride;
    a.chan_off = RDS_HC;
    a.mixI = r->mixI.p;
    a.mixQ = r->mixQ.p;
    a.mix_stride = r->mix_stride;
    a.mix_off = RDS_HM;
    a.ncoI = r->keep_nco ? r->ncoI.p : nullptr;
    a.ncoQ = r->keep_nco ? r->ncoQ.p : nullptr;
    a.nco_stride = r->nco_stride;
    a.n = n;
    a.ncoScale = 0.5;
    a.phaseAdjust = 3 * M_PI / 8;
    dim3 grid((n + 1 + 255) / 256, B);
    sdr_prof_begin(p, "k_rds_mix", s);
    k_rds_mix<<<grid, 256, 0, s>>>(a);
    if ((rc = sdr_check_launch(p, "k_rds_mix"))) return rc;
  }
  {  // R4
    RdsResampleArgs a{};
    a.mixI = r->mixI.p;
    a.mixQ = r->mixQ.p;
    a.mix_stride = r->mix_stride;
    a.mix_off = RDS_HM;
    a.rsI = r->rsI.p;
    a.rsQ = r->rsQ.p;
    a.rs_stride = r->rs_stride;
    a.rs_off = RDS_HS;
    a.quad = r->quad.p;
    a.U = r->U;
    a.D = r->D;
    a.n_out = n_out;
    const int rows_cap = (int)(((long long)(RDS_RS_J - 1) * r->D) / r->U) + 2 + RDS_TP;
    const size_t smem = ((size_t)2 * rows_cap * RDS_RS_PITCH + (size_t)2 * 32 * (RDS_RS_J + 1) + 2 +
                         (size_t)4 * r->quad_rows * 4) * sizeof(double);
    dim3 grid((n_out + RDS_RS_J - 1) / RDS_RS_J, (B + 31) / 32);
    sdr_prof_begin(p, "k_rds_resample", s);
    if (smem > 200 * 1024) return fail(SDR_ERR_INVALID, "RDS resampler tile does not fit in shared memory");
    k_rds_resample<<<grid, 256, smem, s>>>(a, B, rows_cap, n, r->quad_rows);
    if ((rc = sdr_check_launch(p, "k_rds_resample"))) return rc;
  }
  {  // R5
    RdsFirArgs a{};
    a.src[0] = r->rsI.p;
    a.src[1] = r->rsQ.p;
    a.src_stride = r->rs_stride;
    a.src_off = RDS_HS;
    a.dst[0] = r->rrcI.p;
    a.dst[1] = r->rrcQ.p;
    a.dst_stride = r->rrc_stride;
    a.dst_off = 0;
    a.n = n_out;
    dim3 grid(segs_for(n_out, TILE, B, 2, &ops), B, 2);
    a.outs_per_seg = ops;
    sdr_prof_begin(p, "k_rds_fir_rrc", s);
    if (r->precision) k_rds_fir<RDS_TP, R, NT, 2, float><<<grid, NT, 0, s>>>(a, r->h_rrc32);
    else k_rds_fir<RDS_TP, R, NT, 2><<<grid, NT, 0, s>>>(a, r->h_rrc);
    if ((rc = sdr_check_launch(p, "k_rds_fir_rrc"))) return rc;
  }
  {  // R6
    RdsCdrArgs a{};
    a.rrc = r->rrcI.p;
    a.rrc_stride = r->rrc_stride;
    a.block_out = r->block_out;
    a.n_blocks = n_blocks;
    a.sps = r->sps;
    a.first_block = r->blocks_done;
    a.bits = r->bits.p;
    a.counts = r->counts.p;
    a.blocks_cap = r->blocks_cap;
    a.bits_cap = r->bits_cap;
    a.cursor = r->cursor;
    a.batch = B;
    const int total = B * n_blocks;
    sdr_prof_begin(p, "k_rds_cdr", s);
    if (r->cdr_carry)
      k_rds_cdr_carry<<<(B + RDS_CDR_WARPS - 1) / RDS_CDR_WARPS, 32 * RDS_CDR_WARPS, 0, s>>>(a, r->cdr_state.p);
    else
      k_rds_cdr<<<(total + RDS_CDR_WARPS - 1) / RDS_CDR_WARPS, 32 * RDS_CDR_WARPS, 0, s>>>(a);
    if ((rc = sdr_check_launch(p, "k_rds_cdr"))) return rc;
  }
  {  // R7
    RdsCarryArgs c{};
    double *rows[6] = {r->chan.p, r->mixI.p, r->mixQ.p, r->rsI.p, r->rsQ.p, r->theta.p};
    const size_t strides[6] = {r->chan_stride, r->mix_stride, r->mix_stride, r->rs_stride, r->rs_stride,
                               r->theta_stride};
    const int lens[6] = {RDS_HC, RDS_HM, RDS_HM, RDS_HS, RDS_HS, 1};
    const int srcs[6] = {n, n, n, n_out, n_out, n};
    for (int i = 0; i < 6; ++i) {
      c.rows[i] = rows[i];
      c.strides[i] = strides[i];
      c.len[i] = lens[i];
      c.src_off[i] = srcs[i];
    }
    c.demod = r->view.demod;
    c.demod_stride = r->view.stride;
    c.demod_off = r->view.off;
    c.n_if = n;
    c.hist32 = r->hist32.p;
    c.hist_len = RDS_T - 1;
    sdr_prof_begin(p, "k_rds_carry", s);
    k_rds_carry<<<B, 128, 0, s>>>(c);
    if ((rc = sdr_check_launch(p, "k_rds_carry"))) return rc;
  }
  r->cursor += n_blocks;
  r->blocks_done += n_blocks;
  r->last_n_if = n_if;
  r->last_stream = s;
  r->mirrored = false;
  return SDR_OK;
}

static int rds_hook(void *ctx, int event, size_t n_if, cudaStream_t s) {
  sdr_rds *r = static_cast<sdr_rds *>(ctx);
  if (event == 1) return rds_reset_device(r);
  if (event == 2) return rds_validate(r, n_if);
  if (event == 4) {
    if (n_if % (size_t)r->block_if) return fail(SDR_ERR_INVALID, "RDS: the call is not a whole number of RDS blocks");
    if (r->cursor + (long long)(n_if / (size_t)r->block_if) > r->blocks_cap)
      return fail(SDR_ERR_CAPACITY, "RDS: result buffer too small for this call; raise max_pending_blocks or "
                                    "call sdr_rds_read, then sdr_rds_discard");
    return SDR_OK;
  }
  return rds_process(r, n_if, s);
}

// ---------------------------------------------------------------------------
// host bit layer: differential decoding and the frame synchroniser
// ---------------------------------------------------------------------------
// fmSupportLib.py:32-57, one 10-bit mask per row (bit 9 = first column)
static const uint16_t kParityRows[26] = {
    0x200, 0x100, 0x080, 0x040, 0x020, 0x010, 0x008, 0x004, 0x002, 0x001, 0x2DC, 0x16E, 0x0B7,
    0x287, 0x39F, 0x313, 0x355, 0x376, 0x1BB, 0x201, 0x3DC, 0x1EE, 0x0F7, 0x2A7, 0x38F, 0x31B};

// fmSupportLib.py:14-27 and :62-91
static char rds_offset_of(const uint8_t *d) {
  uint16_t s = 0;
  for (int i = 0; i < 26; ++i)
    if (d[i] == 1) s ^= kParityRows[i];
  switch (s) {
    case 0x3D8: return 'A';  // 1111011000
    case 0x3D4: return 'B';  // 1111010100
    case 0x25C: return 'C';  // 1001011100
    case 0x3CC: return 'c';  // 1111001100 (C')
    case 0x258: return 'D';  // 1001011000
  }
  return ' ';
}

// fmSupportLib.py:30-100
static char rds_framesync(const std::vector<uint8_t> &d, size_t *consumed) {
  const long n = (long)d.size();
  long pos = 0;
  char type = ' ';
  while (pos < n - 26) {
    const char o = rds_offset_of(d.data() + pos);
    if (o != ' ') {
      type = o;
      if (n - (pos + 26) < 26) break;
      pos += 26;
    } else {
      pos += 1;
    }
  }
  const long idx = type == ' ' ? pos : pos + 26;
  *consumed = (size_t)std::min(std::max(idx, 0L), n);
  return type;
}

// Bring the pending blocks to the host and run the bit layer for the ones not seen yet.
static int rds_mirror(sdr_rds *r) {
  if (r->mirrored) return SDR_OK;
  SDR_CUDA(cudaSetDevice(r->view.device));
  SDR_CUDA(cudaStreamSynchronize(r->last_stream));
  SDR_CUDA(cudaMemcpy(r->h_bits.data(), r->bits.p, r->h_bits.size(), cudaMemcpyDeviceToHost));
  SDR_CUDA(cudaMemcpy(r->h_counts.data(), r->counts.p, r->h_counts.size() * sizeof(int), cudaMemcpyDeviceToHost));
  const int B = r->view.batch;
  for (int b = 0; b < B; ++b) {
    std::string &offs = r->offsets[(size_t)b];
    for (int blk = (int)offs.size(); blk < r->cursor; ++blk) {
      const int cnt = r->h_counts[(size_t)b * r->blocks_cap + blk];
      const uint8_t *bits = r->h_bits.data() + ((size_t)b * r->blocks_cap + blk) * r->bits_cap;
      std::vector<uint8_t> &dec = r->decoded[(size_t)b];
      std::vector<uint8_t> &df = r->diff[(size_t)b];
      // fmSupportLib.py:241-249: the first bit of a block is taken as it is
      for (int i = 0; i < cnt; ++i) {
        const uint8_t v = i == 0 ? bits[0] : (uint8_t)(bits[i] != bits[i - 1]);
        dec.push_back(v);
        df.push_back(v);
      }
      size_t used = 0;
      offs.push_back(rds_framesync(dec, &used));  // fmRDS.py:275
      dec.erase(dec.begin(), dec.begin() + (long)used);  // fmRDS.py:276
    }
  }
  r->mirrored = true;
  return SDR_OK;
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" int sdr_rds_create(sdr_pipeline *p, const sdr_rds_config *cfg, sdr_rds **out) {
  if (!p || !out) return fail(SDR_ERR_INVALID, "null argument");
  *out = nullptr;
  sdr_rds *r = new (std::nothrow) sdr_rds();
  if (!r) return fail(SDR_ERR_NOMEM, "out of host memory");
  int rc = pipeline_view(p, &r->view);
  if (rc) { delete r; return rc; }
  r->pipe = p;
  r->mode = r->view.mode;
  // fmRDS.py:55-75: RDS is defined for modes 0 and 2 only
  if (r->mode == 0) { r->U = 247; r->D = 960; r->sps = 26; }
  else if (r->mode == 2) { r->U = 817; r->D = 1920; r->sps = 43; }
  else { delete r; return fail(SDR_ERR_INVALID, "RDS is defined for modes 0 and 2 only (fmRDS.py:55-75)"); }
  // fmRDS.py:149-152: block = 2*10*5*960*2 bytes (mode 0) / 10*800*1920*2 bytes (mode 2)
  const int model_block_if = r->mode == 0 ? 9600 : 1536000;
  r->block_if = (cfg && cfg->block_if > 0) ? cfg->block_if : model_block_if;
  // whole resampler outputs and whole audio granules per block
  if (r->block_if % 9600) { delete r; return fail(SDR_ERR_INVALID, "RDS block must be a multiple of 9600 IF samples"); }
  r->block_out = (int)((long long)r->block_if * r->U / r->D);
  r->block_bytes = r->block_if * r->view.rf_decim * 2;
  if ((size_t)r->block_if > r->view.cap_if) {
    delete r;
    return fail(SDR_ERR_CAPACITY, "pipeline max_bytes_per_channel is smaller than one RDS block");
  }
  if (r->block_out <= RDS_CDR_START) { delete r; return fail(SDR_ERR_INVALID, "RDS block too short for the CDR"); }
  {
    const cudaError_t e = cudaSetDevice(r->view.device);
    if (e != cudaSuccess) { delete r; return cuda_fail(e, "cudaSetDevice", __FILE__, __LINE__); }
  }
  // the resampler's tile is the only kernel of the chain above the default 48 KB of dynamic shared memory
  cudaFuncSetAttribute(k_rds_resample, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const size_t B = (size_t)r->view.batch;
  r->cap_if = r->view.cap_if / (size_t)r->block_if * (size_t)r->block_if;
  r->cap_out = r->cap_if * (size_t)r->U / (size_t)r->D;
  const int blocks_per_call = (int)(r->cap_if / (size_t)r->block_if);
  r->blocks_cap = (cfg && cfg->max_pending_blocks > 0) ? cfg->max_pending_blocks : std::max(16, 2 * blocks_per_call);
  r->blocks_cap = std::max(r->blocks_cap, blocks_per_call);
  r->bits_cap = r->block_out / r->sps + 4;  // >= prefix bits + pairs, also with a carried start of 0
  r->keep_nco = cfg && cfg->keep_nco;
  r->cdr_carry = cfg && cfg->cdr_carry;
  r->pll_form = cfg ? cfg->pll_form : 0;
  r->precision = cfg ? cfg->precision : 0;
  if (r->precision < 0 || r->precision > 1) { delete r; return fail(SDR_ERR_INVALID, "unknown sdr_rds_config.precision"); }
  if (r->pll_form < 0 || r->pll_form > 2) { delete r; return fail(SDR_ERR_INVALID, "unknown sdr_rds_config.pll_form"); }
  auto up = [](size_t v) { return (v + 3) / 4 * 4; };
  r->chan_stride = up(RDS_HC + r->cap_if);
  r->carr_stride = up(r->cap_if);
  r->theta_stride = up(r->cap_if + 1);
  r->mix_stride = up(RDS_HM + r->cap_if);
  r->rs_stride = up(RDS_HS + r->cap_out);
  r->rrc_stride = up(r->cap_out);
  r->nco_stride = up(r->cap_if + 1);
  // filters (fmRDS.py:122-125)
  const double if_Fs = (double)r->view.if_Fs;
  std::vector<double> hc = rds_band_pass(RDS_T, if_Fs, 54e3, 60e3);
  std::vector<double> hk = rds_band_pass(RDS_T, if_Fs, 113.5e3, 114.5e3);
  std::vector<double> hr = rds_low_pass(RDS_TP * r->U, if_Fs * r->U, 3e3);
  std::vector<double> hq = rds_rrc(2375.0 * r->sps, RDS_TP);
  std::memcpy(r->h_chan.h, hc.data(), sizeof(r->h_chan.h));
  std::memcpy(r->h_carr.h, hk.data(), sizeof(r->h_carr.h));
  std::memcpy(r->h_rrc.h, hq.data(), sizeof(r->h_rrc.h));
  for (int i = 0; i < RDS_T; ++i) {
    r->h_chan32.h[i] = (float)hc[i];
    r->h_carr32.h[i] = (float)hk[i];
  }
  for (int i = 0; i < RDS_TP; ++i) r->h_rrc32.h[i] = (float)hq[i];
  // Quad table of the resampler: for a quad whose first output has phase p0, output i has
  // phase (p0 + i*D) mod U and its newest input lies floor((p0 + i*D)/U) samples after that of
  // output 0; row r of the entry is the sample r steps before the newest input of output 3.
  r->quad_rows = RDS_TP + (r->U - 1 + 3 * r->D) / r->U;
  std::vector<double> quad((size_t)r->U * r->quad_rows * 4, 0.0);
  for (int p0 = 0; p0 < r->U; ++p0) {
    const int span = (p0 + 3 * r->D) / r->U;
    for (int i = 0; i < 4; ++i) {
      const int ph = (p0 + i * r->D) % r->U, off = (p0 + i * r->D) / r->U;
      for (int k = 0; k < RDS_TP; ++k)
        quad[((size_t)p0 * r->quad_rows + (size_t)(k + span - off)) * 4 + i] = hr[(size_t)ph + (size_t)k * r->U];
    }
  }
  bool ok = !r->quad.alloc(quad.size()) && !r->chan.alloc(B * r->chan_stride) &&
            !r->carr.alloc(B * r->carr_stride) && !r->theta.alloc(B * r->theta_stride) && !r->mixI.alloc(B * r->mix_stride) &&
            !r->mixQ.alloc(B * r->mix_stride) && !r->rsI.alloc(B * r->rs_stride) &&
            !r->rsQ.alloc(B * r->rs_stride) && !r->rrcI.alloc(B * r->rrc_stride) &&
            !r->rrcQ.alloc(B * r->rrc_stride) && !r->pll.alloc(B * 8) && !r->cdr_state.alloc(B * 4) &&
            !r->hist32.alloc(B * (RDS_T - 1)) &&
            !r->bits.alloc(B * (size_t)r->blocks_cap * (size_t)r->bits_cap) &&
            !r->counts.alloc(B * (size_t)r->blocks_cap);
  if (ok && r->keep_nco) ok = !r->ncoI.alloc(B * r->nco_stride) && !r->ncoQ.alloc(B * r->nco_stride);
  if (!ok) { delete r; return SDR_ERR_NOMEM; }
  if (cudaMemcpy(r->quad.p, quad.data(), quad.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
    delete r;
    return fail(SDR_ERR_CUDA, "RDS: uploading the resampler table failed");
  }
  r->h_bits.assign(B * (size_t)r->blocks_cap * (size_t)r->bits_cap, 0);
  r->h_counts.assign(B * (size_t)r->blocks_cap, 0);
  r->decoded.assign(B, {});
  r->offsets.assign(B, {});
  r->diff.assign(B, {});
  if ((rc = rds_reset_device(r)) || (rc = pipeline_set_hook(p, rds_hook, r, r->block_bytes))) {
    delete r;
    return rc;
  }
  *out = r;
  return SDR_OK;
}

extern "C" int sdr_rds_destroy(sdr_rds *r) {
  if (!r) return SDR_OK;
  cudaSetDevice(r->view.device);
  cudaDeviceSynchronize();
  pipeline_set_hook(r->pipe, nullptr, nullptr, 0);
  delete r;
  return SDR_OK;
}

extern "C" int sdr_rds_info(const sdr_rds *r, sdr_rds_info_t *out) {
  if (!r || !out) return fail(SDR_ERR_INVALID, "null argument");
  out->upsamp = r->U;
  out->decim = r->D;
  out->samples_per_symbol = r->sps;
  out->block_if = r->block_if;
  out->block_out = r->block_out;
  out->block_bytes = r->block_bytes;
  out->max_pending_blocks = r->blocks_cap;
  out->max_bits_per_block = r->bits_cap;
  return SDR_OK;
}

extern "C" int sdr_rds_pending(sdr_rds *r, size_t *n_blocks) {
  if (!r || !n_blocks) return fail(SDR_ERR_INVALID, "null argument");
  *n_blocks = (size_t)r->cursor;
  return SDR_OK;
}

extern "C" int sdr_rds_read(sdr_rds *r, int channel, uint8_t *cdr_bits, uint8_t *diff_bits, size_t bits_cap,
                            size_t *n_bits, int *bit_counts, char *offsets, size_t blocks_cap,
                            size_t *n_blocks) {
  if (!r) return fail(SDR_ERR_INVALID, "null argument");
  if (channel < 0 || channel >= r->view.batch) return fail(SDR_ERR_INVALID, "channel out of range");
  int rc = rds_mirror(r);
  if (rc) return rc;
  const size_t nb = (size_t)r->cursor;
  size_t total = 0;
  for (size_t blk = 0; blk < nb; ++blk) total += (size_t)r->h_counts[(size_t)channel * r->blocks_cap + blk];
  if (n_bits) *n_bits = total;
  if (n_blocks) *n_blocks = nb;
  if ((cdr_bits || diff_bits) && bits_cap < total) return fail(SDR_ERR_CAPACITY, "bit buffer too small");
  if ((bit_counts || offsets) && blocks_cap < nb) return fail(SDR_ERR_CAPACITY, "block buffer too small");
  size_t w = 0;
  for (size_t blk = 0; blk < nb; ++blk) {
    const int cnt = r->h_counts[(size_t)channel * r->blocks_cap + blk];
    if (cdr_bits)
      std::memcpy(cdr_bits + w, r->h_bits.data() + ((size_t)channel * r->blocks_cap + blk) * r->bits_cap, (size_t)cnt);
    if (bit_counts) bit_counts[blk] = cnt;
    w += (size_t)cnt;
  }
  if (diff_bits && total) std::memcpy(diff_bits, r->diff[(size_t)channel].data(), total);
  if (offsets && nb) std::memcpy(offsets, r->offsets[(size_t)channel].data(), nb);
  return SDR_OK;
}

extern "C" int sdr_rds_discard(sdr_rds *r) {
  if (!r) return fail(SDR_ERR_INVALID, "null argument");
  int rc = rds_mirror(r);  // the frame synchroniser still has to see every block
  if (rc) return rc;
  r->cursor = 0;
  for (auto &o : r->offsets) o.clear();
  for (auto &d : r->diff) d.clear();
  return SDR_OK;
}

extern "C" int sdr_rds_tap(sdr_rds *r, int stage, int channel, double *dst, size_t cap, size_t *n) {
  if (!r || !n) return fail(SDR_ERR_INVALID, "null argument");
  if (channel < 0 || channel >= r->view.batch) return fail(SDR_ERR_INVALID, "channel out of range");
  const size_t n_if = r->last_n_if;
  const size_t n_out = n_if * (size_t)r->U / (size_t)r->D;
  const double *src = nullptr;
  size_t cnt = 0;
  // The carry has already moved the tails into the prefixes; samples 0.. of the last call are
  // still in place behind them.
  switch (stage) {
    case SDR_RDS_TAP_CHANNEL: src = r->chan.p + (size_t)channel * r->chan_stride + RDS_HC; cnt = n_if; break;
    case SDR_RDS_TAP_CARRIER: src = r->carr.p + (size_t)channel * r->carr_stride; cnt = n_if; break;
    case SDR_RDS_TAP_PLL_I: src = r->ncoI.p ? r->ncoI.p + (size_t)channel * r->nco_stride : nullptr; cnt = n_if + 1; break;
    case SDR_RDS_TAP_PLL_Q: src = r->ncoQ.p ? r->ncoQ.p + (size_t)channel * r->nco_stride : nullptr; cnt = n_if + 1; break;
    case SDR_RDS_TAP_MIXER_I: src = r->mixI.p + (size_t)channel * r->mix_stride + RDS_HM; cnt = n_if; break;
    case SDR_RDS_TAP_MIXER_Q: src = r->mixQ.p + (size_t)channel * r->mix_stride + RDS_HM; cnt = n_if; break;
    case SDR_RDS_TAP_RESAMPLER_I: src = r->rsI.p + (size_t)channel * r->rs_stride + RDS_HS; cnt = n_out; break;
    case SDR_RDS_TAP_RESAMPLER_Q: src = r->rsQ.p + (size_t)channel * r->rs_stride + RDS_HS; cnt = n_out; break;
    case SDR_RDS_TAP_RRC_I: src = r->rrcI.p + (size_t)channel * r->rrc_stride; cnt = n_out; break;
    case SDR_RDS_TAP_RRC_Q: src = r->rrcQ.p + (size_t)channel * r->rrc_stride; cnt = n_out; break;
    default: return fail(SDR_ERR_INVALID, "unknown RDS stage");
  }
  if (!src) return fail(SDR_ERR_INVALID, "NCO outputs are kept only when sdr_rds_config.keep_nco is set");
  if (n_if == 0) cnt = 0;
  *n = cnt;
  if (!dst) return SDR_OK;
  if (cap < cnt) return fail(SDR_ERR_CAPACITY, "destination too small");
  SDR_CUDA(cudaSetDevice(r->view.device));
  SDR_CUDA(cudaStreamSynchronize(r->last_stream));
  if (cnt) SDR_CUDA(cudaMemcpy(dst, src, cnt * sizeof(double), cudaMemcpyDeviceToHost));
  return SDR_OK;
}

extern "C" int sdr_rds_design(int which, int mode, double *h, size_t cap, size_t *n) {
  if (!n) return fail(SDR_ERR_INVALID, "null argument");
  if (mode != 0 && mode != 2) return fail(SDR_ERR_INVALID, "RDS is defined for modes 0 and 2 only");
  const int U = mode == 0 ? 247 : 817, sps = mode == 0 ? 26 : 43;
  std::vector<double> v;
  switch (which) {
    case SDR_RDS_FILTER_CHANNEL: v = rds_band_pass(RDS_T, 240000.0, 54e3, 60e3); break;
    case SDR_RDS_FILTER_CARRIER: v = rds_band_pass(RDS_T, 240000.0, 113.5e3, 114.5e3); break;
    case SDR_RDS_FILTER_RESAMPLER: v = rds_low_pass(RDS_TP * U, 240000.0 * U, 3e3); break;
    case SDR_RDS_FILTER_RRC: v = rds_rrc(2375.0 * sps, RDS_TP); break;
    default: return fail(SDR_ERR_INVALID, "unknown RDS filter");
  }
  *n = v.size();
  if (!h) return SDR_OK;
  if (cap < v.size()) return fail(SDR_ERR_CAPACITY, "destination too small");
  std::memcpy(h, v.data(), v.size() * sizeof(double));
  return SDR_OK;
}
