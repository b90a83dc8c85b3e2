// kernels.cuh -- sm_100a kernels of the FM receiver DSP path (exact variant).
//
// Layout in HBM (all channel-major, one row per capture):
//   iq      [B][iq_stride]            uint8, interleaved I,Q (caller's buffer)
//   rf_hist [B][2*HR]                 uint8, last HR I/Q pairs of the previous call
//   demod   [B][HD + n_if (+pad)]     float, HD-sample history prefix then this call's samples
//   stf     [B][HA + n_if (+pad)]     float, stereo band-pass output, history-prefixed
//   car     [B][n_if]                 float, pilot band-pass output
//   nco     [B][HA + n_if + 1 (+pad)] float, PLL output delayed by one sample (what the
//                                     reference's mixer indexes, project.cpp:246-248)
//   pcm     [B][pcm_stride]           int16 (mono) or L,R interleaved (stereo)
// A history prefix plays the role of the reference's per-filter `state` vectors
// (project.cpp:29-36,61-65): the last samples of the previous call are kept in
// front of the current ones, so every FIR reads one contiguous window and no
// kernel branches on "previous block or this block" (filter.cpp:141-145).
//
// Every FIR accumulates exactly like the reference: float accumulator from +0,
// taps in ascending order, separately rounded multiply and add.
#pragma once

#include "common.cuh"
#include "libm_exact.cuh"

namespace sdr {

constexpr int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Taps travel as a kernel parameter: with the tap index a compile-time constant
// after unrolling, every tap is a constant-bank operand of its FMUL and costs no
// load instruction and no register.
template <int N>
struct TapArray {
  float h[N];
};
// Tap storage needed by the two FIR cores for a T-tap, decimate-by-D filter.
constexpr int taps_window(int T) { return round_up(T, 4); }
constexpr int fir_groups_count(int T, int D, int R) { return round_up((T + D - 1) / D, R); }
constexpr int taps_groups(int T, int D, int R) { return fir_groups_count(T, D, R) * D; }

// ---------------------------------------------------------------------------
// Register-tiled FIR core.  One thread produces R consecutive outputs of a
// decimate-by-D, T-tap FIR from a window in shared memory.  `w` points at the
// sample with index (first_output*D - HALO) and must be 16-byte aligned; the
// window is walked from the newest sample to the oldest so that each accumulator
// sees its taps in ascending order (n = r*D - e grows as e falls).
// ---------------------------------------------------------------------------
// Shared-memory rows are stored in groups of G = R*D samples (one thread's share of a
// tile) separated by PAD floats, chosen so that the per-thread stride (G+PAD)/4 is odd:
// a quarter-warp's LDS.128 then touches 8 distinct 16-byte bank groups (no conflict).
constexpr int fir_pad(int G) { return ((G / 4) % 2 == 1) ? 0 : 4; }
constexpr int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
// Offset (in floats) of sample `e` relative to the first sample of a thread's own group.
constexpr int fir_off(int e, int G) { return e + fir_pad(G) * floordiv(e, G); }

// FMA = true contracts each multiply-add (SDR_VARIANT_FAST only: no longer bit-identical).
template <int T, int D, int R, int HALO, bool FMA = false>
__device__ __forceinline__ void fir_window(const float *__restrict__ w,
                                           const TapArray<taps_window(T)> &taps, float (&acc)[R]) {
  // `w` points at the sample of the thread's FIRST output (e = 0).
  static_assert(HALO % 4 == 0 && HALO >= T - 1, "halo must cover the filter and be float4-aligned");
  static_assert((R * D) % 4 == 0, "group size must keep float4 alignment");
  constexpr int G = R * D;
  constexpr int NEWEST = (R - 1) * D;              // newest sample used, relative to first output
  constexpr int C_HI = (HALO + NEWEST) / 4;        // chunk holding the newest sample
  constexpr int C_LO = (HALO - (T - 1)) / 4;       // chunk holding the oldest sample
#pragma unroll
  for (int c = C_HI; c >= C_LO; --c) {
    const float4 v = *reinterpret_cast<const float4 *>(w + fir_off(4 * c - HALO, G));
#pragma unroll
    for (int j = 3; j >= 0; --j) {
      const int e = 4 * c + j - HALO;
      const float xv = (j == 0) ? v.x : (j == 1) ? v.y : (j == 2) ? v.z : v.w;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int n = r * D - e;
        if (n >= 0 && n < T) acc[r] = FMA ? __fmaf_rn(taps.h[n], xv, acc[r]) : xmac(acc[r], taps.h[n], xv);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Looped FIR core ("tap groups").  The T taps are cut into groups of D:
//   y_r = sum_g sum_j h[D*g + j] * x[D*(r-g) - j]
// so that in step g every one of the thread's R outputs uses the SAME D taps (which
// therefore live in uniform registers / the constant bank) and the input block
// b = r-g of D samples.  Going from g to g+1 the R blocks in flight shift by one:
// R-1 stay in registers, one new block (b = -g) is loaded.  Unrolling the g loop by
// R makes the register renaming static, the loop body stays R*R*D multiply-adds
// however long the filter is -- a few KB of code instead of the tens of KB of the
// fully unrolled window walk, which is what the instruction cache needs (profiles/).
// Per output the taps are still visited in ascending order n = D*g + j with one
// rounding per multiply and per add, i.e. bit-identical to the reference loops.
// The tap array is zero-padded to a whole number of groups; a zero tap adds +-0 to
// an accumulator that is never -0, so padding changes no bit as long as the inputs
// are finite (true for every signal inside the pipeline).
// `w` points at the thread's first output's newest sample (e = 0) inside a padded
// row (RowGeom); one thread's share of a tile is exactly one group of G = R*D floats.
// ---------------------------------------------------------------------------
template <int T, int D, int R>
__device__ __forceinline__ void fir_groups(const float *__restrict__ w,
                                           const TapArray<taps_groups(T, D, R)> &taps,
                                           float (&acc)[R]) {
  constexpr int G = R * D;
  constexpr int STRIDE = G + fir_pad(G);
  constexpr int NG = fir_groups_count(T, D, R);
  static_assert(G % 4 == 0, "group must keep float4 alignment");
  float xb[R][D];  // block in slot s holds x[D*b - j], j = 0..D-1, for the block b == s (mod R)
  // prologue: blocks 1..R-1 lie in the thread's own group, e in [1, D*(R-1)]
  {
    constexpr int NCH = (D * (R - 1)) / 4 + 1;
    float own[NCH * 4];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const float4 v = *reinterpret_cast<const float4 *>(w + 4 * c);
      own[4 * c] = v.x; own[4 * c + 1] = v.y; own[4 * c + 2] = v.z; own[4 * c + 3] = v.w;
    }
#pragma unroll
    for (int b = 1; b < R; ++b)
#pragma unroll
      for (int j = 0; j < D; ++j) xb[b][j] = own[D * b - j];
  }
#pragma unroll 1
  for (int p = 0; p < NG / R; ++p) {
    const float *base = w - p * STRIDE;   // e = -G*p
    const float *hp = taps.h + p * G;     // taps of groups R*p .. R*p+R-1
    // samples e_rel in [-G, 3]: the previous group (one STRIDE back) plus the first chunk here
    float xv[G + 4];
#pragma unroll
    for (int c = 0; c < G / 4; ++c) {
      const float4 v = *reinterpret_cast<const float4 *>(base - STRIDE + 4 * c);
      xv[4 * c] = v.x; xv[4 * c + 1] = v.y; xv[4 * c + 2] = v.z; xv[4 * c + 3] = v.w;
    }
    {
      const float4 v = *reinterpret_cast<const float4 *>(base);
      xv[G] = v.x; xv[G + 1] = v.y; xv[G + 2] = v.z; xv[G + 3] = v.w;
    }
#pragma unroll
    for (int c = 0; c < R; ++c) {
      // new block b = -(R*p + c): e_rel = -D*c - j  ->  xv[G - D*c - j]
#pragma unroll
      for (int j = 0; j < D; ++j) xb[(R - c) % R][j] = xv[G - D * c - j];
#pragma unroll
      for (int j = 0; j < D; ++j)
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = xmac(acc[r], hp[D * c + j], xb[(r - c + R) % R][j]);
    }
  }
}

// Two filters over the same input window (the stereo and pilot band-pass filters read
// the same demodulated samples): identical walk, two tap sets, two accumulator sets.
// FMA_A = true contracts the multiply-adds of filter A only (SDR_VARIANT_MIXED: the stereo band
// does not feed the PLL); filter B stays in the reference's two-rounding form.
template <int T, int D, int R, bool FMA_A = false>
__device__ __forceinline__ void fir_groups2(const float *__restrict__ w,
                                            const TapArray<taps_groups(T, D, R)> &taps_a,
                                            const TapArray<taps_groups(T, D, R)> &taps_b,
                                            float (&acc_a)[R], float (&acc_b)[R]) {
  constexpr int G = R * D;
  constexpr int STRIDE = G + fir_pad(G);
  constexpr int NG = fir_groups_count(T, D, R);
  static_assert(G % 4 == 0, "group must keep float4 alignment");
  float xb[R][D];
  {
    constexpr int NCH = (D * (R - 1)) / 4 + 1;
    float own[NCH * 4];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const float4 v = *reinterpret_cast<const float4 *>(w + 4 * c);
      own[4 * c] = v.x; own[4 * c + 1] = v.y; own[4 * c + 2] = v.z; own[4 * c + 3] = v.w;
    }
#pragma unroll
    for (int b = 1; b < R; ++b)
#pragma unroll
      for (int j = 0; j < D; ++j) xb[b][j] = own[D * b - j];
  }
#pragma unroll 1
  for (int p = 0; p < NG / R; ++p) {
    const float *base = w - p * STRIDE;
    const float *ha = taps_a.h + p * G;
    const float *hb = taps_b.h + p * G;
    float xv[G + 4];
#pragma unroll
    for (int c = 0; c < G / 4; ++c) {
      const float4 v = *reinterpret_cast<const float4 *>(base - STRIDE + 4 * c);
      xv[4 * c] = v.x; xv[4 * c + 1] = v.y; xv[4 * c + 2] = v.z; xv[4 * c + 3] = v.w;
    }
    {
      const float4 v = *reinterpret_cast<const float4 *>(base);
      xv[G] = v.x; xv[G + 1] = v.y; xv[G + 2] = v.z; xv[G + 3] = v.w;
    }
#pragma unroll
    for (int c = 0; c < R; ++c) {
#pragma unroll
      for (int j = 0; j < D; ++j) xb[(R - c) % R][j] = xv[G - D * c - j];
#pragma unroll
      for (int j = 0; j < D; ++j)
#pragma unroll
        for (int r = 0; r < R; ++r) {
          acc_a[r] = FMA_A ? __fmaf_rn(ha[D * c + j], xb[(r - c + R) % R][j], acc_a[r])
                           : xmac(acc_a[r], ha[D * c + j], xb[(r - c + R) % R][j]);
          acc_b[r] = xmac(acc_b[r], hb[D * c + j], xb[(r - c + R) % R][j]);
        }
    }
  }
}

// Packed form of fir_groups2 (both filters exact): accumulator pair (A_r, B_r), tap pair
// (hA[n], hB[n]) from the constant bank, the input sample broadcast to both lanes; one
// FMUL2 + FFMA2 per tap and output where the scalar form issues four instructions.
template <int N>
struct TapPairs {
  float2 h[N];
};
template <int T, int D, int R>
__device__ __forceinline__ void fir_groups2_packed(const float *__restrict__ w,
                                                   const TapPairs<taps_groups(T, D, R)> &taps,
                                                   f32x2_t (&acc)[R], f32x2_t one) {
  constexpr int G = R * D;
  constexpr int STRIDE = G + fir_pad(G);
  constexpr int NG = fir_groups_count(T, D, R);
  static_assert(G % 4 == 0, "group must keep float4 alignment");
  float xb[R][D];
  {
    constexpr int NCH = (D * (R - 1)) / 4 + 1;
    float own[NCH * 4];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const float4 v = *reinterpret_cast<const float4 *>(w + 4 * c);
      own[4 * c] = v.x; own[4 * c + 1] = v.y; own[4 * c + 2] = v.z; own[4 * c + 3] = v.w;
    }
#pragma unroll
    for (int b = 1; b < R; ++b)
#pragma unroll
      for (int j = 0; j < D; ++j) xb[b][j] = own[D * b - j];
  }
#pragma unroll 1
  for (int p = 0; p < NG / R; ++p) {
    const float *base = w - p * STRIDE;
    const float2 *hp = taps.h + p * G;
    float xv[G + 4];
#pragma unroll
    for (int c = 0; c < G / 4; ++c) {
      const float4 v = *reinterpret_cast<const float4 *>(base - STRIDE + 4 * c);
      xv[4 * c] = v.x; xv[4 * c + 1] = v.y; xv[4 * c + 2] = v.z; xv[4 * c + 3] = v.w;
    }
    {
      const float4 v = *reinterpret_cast<const float4 *>(base);
      xv[G] = v.x; xv[G + 1] = v.y; xv[G + 2] = v.z; xv[G + 3] = v.w;
    }
#pragma unroll
    for (int c = 0; c < R; ++c) {
#pragma unroll
      for (int j = 0; j < D; ++j) xb[(R - c) % R][j] = xv[G - D * c - j];
#pragma unroll
      for (int j = 0; j < D; ++j) {
        const f32x2_t h2 = pack2(hp[D * c + j].x, hp[D * c + j].y);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float x = xb[(r - c + R) % R][j];
          acc[r] = xmac2(acc[r], h2, pack2(x, x), one);
        }
      }
    }
  }
}

// Geometry of one padded shared-memory row holding tile samples [-HALO, TILE_IN).
template <int D, int R, int NT, int HALO>
struct RowGeom {
  static constexpr int G = R * D;
  static constexpr int PAD = fir_pad(G);
  static constexpr int HALO_G = (HALO + G - 1) / G;       // groups in front of sample 0
  static constexpr int ORIGIN = HALO_G * (G + PAD);       // float position of sample 0
  static constexpr int FLOATS = ORIGIN + NT * (G + PAD) + 8;
  // position of tile sample a (a >= -HALO); a may be a run-time value
  __device__ static __forceinline__ int pos(int a) {
    const int ap = a + HALO_G * G;
    return ap + PAD * (ap / G);
  }
  __device__ static __forceinline__ int thread_base(int t) { return ORIGIN + t * (G + PAD); }
};

// ---------------------------------------------------------------------------
// K1: RF front end.  uint8 I/Q -> (u8-128)/128 -> low-pass + decimate (I and Q)
// -> FM discriminator.  Replaces readStdinBlockData (iofunc.cpp:128-135), the
// de-interleave (project.cpp:101-105), two convolveBlockFastFIR calls (:111,:121)
// and fmDemod (:128).
// Grid: (segments, B).  Each CTA walks its segment tile by tile; a tile is
// NT*R outputs = NT*R*D input pairs staged in shared memory as floats.
// ---------------------------------------------------------------------------
struct RfArgs {
  const uint8_t *iq;       // [B][iq_stride]
  size_t iq_stride;        // bytes
  const uint8_t *hist;     // [B][2*HR]   (HR = rf_hist_len)
  int rf_hist_len;         // HR
  const float *prev_in;    // [B][2]  I,Q of the last output of the previous call
  float *prev_out;         // [B][2]
  float *demod;            // [B][demod_stride], sample 0 at demod_off
  size_t demod_stride;
  int demod_off;
  float *i_filt, *q_filt;  // optional [B][tap_stride]
  size_t tap_stride;
  long long n_rf;          // input pairs per capture in this call
  int n_if;                // outputs per capture in this call
  int outs_per_seg;        // multiple of the tile size
  // tensor-core front end only: fm_demod also (or only) as two half-precision planes
  // x = xh + xl (xh = fp16(x), xl = fp16(x - xh)): the operand format of the tensor-core
  // resampler (resample_tc.cuh).  [B][pl_stride] halfs each, sample 0 at pl_off.
  uint16_t *xh, *xl;
  size_t pl_stride;
  int pl_off;
  int write_f32;           // 0: the planes replace the float row (nobody else reads fm_demod)
  float one;               // 1.0f, known only at run time (xmac2's multiplier; see common.cuh)
};

template <int T, int D>
__device__ __forceinline__ void rf_fetch_pair(const RfArgs &a, const uint8_t *row,
                                              const uint8_t *hrow, long long i, float &fi,
                                              float &fq) {
  uint32_t bi = 128, bq = 128;
  if (i < 0) {
    long long k = a.rf_hist_len + i;
    if (k >= 0) {
      bi = hrow[2 * k];
      bq = hrow[2 * k + 1];
    }
  } else if (i < a.n_rf) {
    bi = row[2 * i];
    bq = row[2 * i + 1];
  }
  fi = u8_centered(bi);
  fq = u8_centered(bq);
}

template <int T, int D, int R, int NT, int ALGO>
struct RfCfg {
  static constexpr int NTAPS = ALGO ? taps_groups(T, D, R) : taps_window(T);
  static constexpr int HALO_MIN = ALGO ? (fir_groups_count(T, D, R) * D + D) : (T - 1 + D);
  static constexpr int HALO = round_up(HALO_MIN, 8);
  static constexpr int TILE_OUT = NT * R;
  static constexpr int TILE_IN = TILE_OUT * D;
  using Geom = RowGeom<D, R, NT, HALO>;
  static constexpr int ROW = round_up(Geom::FLOATS, 4);  // floats per component in smem
  static constexpr size_t SMEM = (size_t)ROW * 2 * sizeof(float);
};

// ALGO 0: fully unrolled window walk (fir_window); ALGO 1: looped tap groups (fir_groups).
template <int T, int D, int R, int NT, bool MERGE, int ALGO>
__global__ void __launch_bounds__(NT)
k_rf_demod(const RfArgs a, const __grid_constant__ TapArray<RfCfg<T, D, R, NT, ALGO>::NTAPS> taps) {
  using Cfg = RfCfg<T, D, R, NT, ALGO>;
  using Geom = typename Cfg::Geom;
  constexpr int HALO = Cfg::HALO;
  extern __shared__ __align__(16) float smem[];
  float *xi = smem;
  float *xq = smem + Cfg::ROW;
  __shared__ float edge_i[NT + 1], edge_q[NT + 1];

  const int t = threadIdx.x;
  const int b = blockIdx.y;
  const int o_begin = blockIdx.x * a.outs_per_seg;
  const int o_end = min(o_begin + a.outs_per_seg, a.n_if);
  if (o_begin >= o_end) return;
  const uint8_t *row = a.iq + (size_t)b * a.iq_stride;
  const uint8_t *hrow = a.hist + (size_t)b * 2 * a.rf_hist_len;
  const bool row_aligned = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);

  for (int o0 = o_begin; o0 < o_end; o0 += Cfg::TILE_OUT) {
    __syncthreads();  // everyone is done with the previous tile's smem and edges
    if (t == 0 && o0 != o_begin) {
      edge_i[0] = edge_i[NT];
      edge_q[0] = edge_q[NT];
    }
    // ---- stage [o0*D - HALO, o0*D + TILE_IN) as centred floats ----
    const long long s0 = (long long)o0 * D - HALO;
    constexpr int NCHUNK = (HALO + Cfg::TILE_IN) / 8;
    for (int q = t; q < NCHUNK; q += NT) {
      const long long i = s0 + 8ll * q;
      float fi[8], fq[8];
      if (row_aligned && i >= 0 && i + 8 <= a.n_rf) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(row + 2 * i));
        const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          fi[2 * k] = u8_centered(wds[k] & 0xffu);
          fq[2 * k] = u8_centered((wds[k] >> 8) & 0xffu);
          fi[2 * k + 1] = u8_centered((wds[k] >> 16) & 0xffu);
          fq[2 * k + 1] = u8_centered(wds[k] >> 24);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) rf_fetch_pair<T, D>(a, row, hrow, i + k, fi[k], fq[k]);
      }
      const int p0 = Geom::pos(8 * q - HALO), p1 = Geom::pos(8 * q + 4 - HALO);
      *reinterpret_cast<float4 *>(xi + p0) = make_float4(fi[0], fi[1], fi[2], fi[3]);
      *reinterpret_cast<float4 *>(xi + p1) = make_float4(fi[4], fi[5], fi[6], fi[7]);
      *reinterpret_cast<float4 *>(xq + p0) = make_float4(fq[0], fq[1], fq[2], fq[3]);
      *reinterpret_cast<float4 *>(xq + p1) = make_float4(fq[4], fq[5], fq[6], fq[7]);
    }
    __syncthreads();
    // ---- I,Q of the output that precedes this segment (fmDemod's prev_i/prev_q) ----
    if (o0 == o_begin && t < 2) {
      float v;
      if (o_begin == 0) {
        v = a.prev_in[2 * b + t];
      } else {
        const float *src = (t == 0 ? xi : xq);
        float acc = 0.0f;  // output o0-1: newest sample is tile sample -D
#pragma unroll 1
        for (int n = 0; n < T; ++n) acc = xmac(acc, taps.h[n], src[Geom::pos(-D - n)]);
        v = xmul(acc, 0.0078125f);
      }
      (t == 0 ? edge_i : edge_q)[0] = v;
    }
    // ---- R outputs per thread for I and Q ----
    float ai[R], aq[R];
    if (MERGE) {
      // one copy of the unrolled FIR body, run twice: halves the instruction footprint
#pragma unroll 1
      for (int comp = 0; comp < 2; ++comp) {
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.0f;
        if constexpr (ALGO) fir_groups<T, D, R>(smem + comp * Cfg::ROW + Geom::thread_base(t), taps, acc);
        else fir_window<T, D, R, HALO>(smem + comp * Cfg::ROW + Geom::thread_base(t), taps, acc);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (comp == 0) ai[r] = acc[r];
          else aq[r] = acc[r];
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) ai[r] = aq[r] = 0.0f;
      if constexpr (ALGO) {
        fir_groups<T, D, R>(xi + Geom::thread_base(t), taps, ai);
        fir_groups<T, D, R>(xq + Geom::thread_base(t), taps, aq);
      } else {
        fir_window<T, D, R, HALO>(xi + Geom::thread_base(t), taps, ai);
        fir_window<T, D, R, HALO>(xq + Geom::thread_base(t), taps, aq);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {  // exact power-of-two scaling: (u8-128)/128
      ai[r] = xmul(ai[r], 0.0078125f);
      aq[r] = xmul(aq[r], 0.0078125f);
    }
    edge_i[t + 1] = ai[R - 1];
    edge_q[t + 1] = aq[R - 1];
    __syncthreads();
    float pi = edge_i[t], pq = edge_q[t];
    const int o = o0 + t * R;
    float *drow = a.demod + (size_t)b * a.demod_stride + a.demod_off;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (o + r < o_end) {
        drow[o + r] = fm_demod_one(ai[r], aq[r], pi, pq);
        if (a.i_filt) {
          a.i_filt[(size_t)b * a.tap_stride + o + r] = ai[r];
          a.q_filt[(size_t)b * a.tap_stride + o + r] = aq[r];
        }
        if (o + r == a.n_if - 1) {
          a.prev_out[2 * b] = ai[r];
          a.prev_out[2 * b + 1] = aq[r];
        }
      }
      pi = ai[r];
      pq = aq[r];
    }
  }
}

// ---------------------------------------------------------------------------
// K1, packed form (the exact variant's default for the 151-tap filter).  I and Q of a sample
// sit side by side in shared memory and in one 64-bit register pair; each tap is applied to both
// with xmac2 (FMUL2 + FFMA2, the tap a broadcast uniform-register operand), so the
// separately rounded multiply-adds of an output's I and Q take 302 issue slots instead of 604
// and the tap loads are shared by the two components.  Per lane the arithmetic is xmac's:
// bit-identical to k_rf_demod (tests run both against the oracle).  Four outputs per thread: the
// window walk at two per thread ran the shared-memory data pipe at 93 % of its wavefront peak
// (each staged pair is re-read by ~8 threads); at four it is 74 %, level with the FP32 pipe.
// ---------------------------------------------------------------------------
template <int D, int R, int NT, int HALO>
struct RowGeomIQ {   // positions in I/Q PAIRS (8 bytes); one LDS.128 = 2 pairs
  static constexpr int G = R * D;
  static_assert(G % 2 == 0 && HALO % 2 == 0, "a 16-byte unit must not straddle two threads' groups");
  static constexpr int PAD = ((G / 2) % 2 == 1) ? 0 : 2;   // per-thread stride in 16-byte units is odd
  static constexpr int HALO_G = (HALO + G - 1) / G;
  static constexpr int ORIGIN = HALO_G * (G + PAD);
  static constexpr int PAIRS = round_up(ORIGIN + NT * (G + PAD) + 4, 2);
  __device__ static __forceinline__ int pos(int a) {
    const int ap = a + HALO_G * G;
    return ap + PAD * (ap / G);
  }
  static constexpr int off(int e) { return e + PAD * floordiv(e, G); }   // relative to a thread's own group
  __device__ static __forceinline__ int thread_base(int t) { return ORIGIN + t * (G + PAD); }
};

// Window walk on PAIRS: R consecutive outputs of a decimate-by-D, T-tap FIR applied to both lanes of
// a staged float2 stream (I and Q of the front end; mono-path and stereo-path input of the stereo audio
// filter), newest sample first so that every accumulator sees its taps in ascending order.  `w` points
// at the pair of the thread's first output's newest sample inside a RowGeomIQ row.
template <int T, int D, int R, int HALO, typename Geom>
__device__ __forceinline__ void fir_window_pairs(const float2 *__restrict__ w, const TapArray<taps_window(T)> &taps,
                                                 f32x2_t (&acc)[R], f32x2_t one) {
  constexpr int NEWEST = (R - 1) * D;
  constexpr int C_HI = (HALO + NEWEST) / 2;
  constexpr int C_LO = (HALO - (T - 1)) / 2;
#pragma unroll
  for (int c = C_HI; c >= C_LO; --c) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(w + Geom::off(2 * c - HALO));
#pragma unroll
    for (int j = 1; j >= 0; --j) {
      const int e = 2 * c + j - HALO;
      const f32x2_t x = j ? v.y : v.x;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int n = r * D - e;
        if (n >= 0 && n < T) acc[r] = xmac2(acc[r], pack2(taps.h[n], taps.h[n]), x, one);
      }
    }
  }
}

template <int T, int D, int R, int NT>
struct RfIqCfg {
  static constexpr int HALO = round_up(T - 1 + D, 8);
  static constexpr int TILE_OUT = NT * R;
  static constexpr int TILE_IN = TILE_OUT * D;
  using Geom = RowGeomIQ<D, R, NT, HALO>;
  static constexpr size_t SMEM = (size_t)Geom::PAIRS * sizeof(float2);
};

template <int T, int D, int R, int NT>
__global__ void __launch_bounds__(NT)
k_rf_demod_iq(const RfArgs a, const __grid_constant__ TapArray<taps_window(T)> taps) {
  using Cfg = RfIqCfg<T, D, R, NT>;
  using Geom = typename Cfg::Geom;
  constexpr int HALO = Cfg::HALO;
  extern __shared__ __align__(16) float2 xiq[];
  __shared__ float edge_i[NT + 1], edge_q[NT + 1];

  const int t = threadIdx.x;
  const int b = blockIdx.y;
  const int o_begin = blockIdx.x * a.outs_per_seg;
  const int o_end = min(o_begin + a.outs_per_seg, a.n_if);
  if (o_begin >= o_end) return;
  const uint8_t *row = a.iq + (size_t)b * a.iq_stride;
  const uint8_t *hrow = a.hist + (size_t)b * 2 * a.rf_hist_len;
  const bool row_aligned = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
  const f32x2_t one = pack2(a.one, a.one);

  // A tile's raw bytes (16-byte chunks of 8 I/Q pairs, NPRE per thread) are loaded into registers one
  // tile ahead: the loads of tile k+1 are in flight while tile k is computed, so only a segment's
  // first tile waits for HBM (without this 40 % of the stall samples sat on the first use of the
  // loaded word, profiles/r2).
  constexpr int NCHUNK = (HALO + Cfg::TILE_IN) / 8;
  constexpr int NPRE = (NCHUNK + NT - 1) / NT;
  uint4 pre[NPRE];
  auto chunk_in_row = [&](long long i) { return row_aligned && i >= 0 && i + 8 <= a.n_rf; };
  auto fetch = [&](int o0) {
    const long long s0 = (long long)o0 * D - HALO;
#pragma unroll
    for (int k = 0; k < NPRE; ++k) {
      const int q = t + k * NT;
      const long long i = s0 + 8ll * q;
      if (q < NCHUNK && chunk_in_row(i)) pre[k] = __ldg(reinterpret_cast<const uint4 *>(row + 2 * i));
    }
  };
  fetch(o_begin);

  for (int o0 = o_begin; o0 < o_end; o0 += Cfg::TILE_OUT) {
    __syncthreads();
    if (t == 0 && o0 != o_begin) {
      edge_i[0] = edge_i[NT];
      edge_q[0] = edge_q[NT];
    }
    // ---- stage [o0*D - HALO, o0*D + TILE_IN) as centred float pairs, in the input's own order ----
    const long long s0 = (long long)o0 * D - HALO;
#pragma unroll
    for (int k = 0; k < NPRE; ++k) {
      const int q = t + k * NT;
      if (q >= NCHUNK) break;
      const long long i = s0 + 8ll * q;
      float fi[8], fq[8];
      if (chunk_in_row(i)) {
        const uint32_t wds[4] = {pre[k].x, pre[k].y, pre[k].z, pre[k].w};
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          fi[2 * m] = u8_centered(wds[m] & 0xffu);
          fq[2 * m] = u8_centered((wds[m] >> 8) & 0xffu);
          fi[2 * m + 1] = u8_centered((wds[m] >> 16) & 0xffu);
          fq[2 * m + 1] = u8_centered(wds[m] >> 24);
        }
      } else {
#pragma unroll
        for (int m = 0; m < 8; ++m) rf_fetch_pair<T, D>(a, row, hrow, i + m, fi[m], fq[m]);
      }
#pragma unroll
      for (int m = 0; m < 8; m += 2)
        *reinterpret_cast<float4 *>(xiq + Geom::pos(8 * q + m - HALO)) = make_float4(fi[m], fq[m], fi[m + 1], fq[m + 1]);
    }
    if (o0 + Cfg::TILE_OUT < o_end) fetch(o0 + Cfg::TILE_OUT);
    __syncthreads();
    // ---- I,Q of the output that precedes this segment (fmDemod's prev_i/prev_q) ----
    if (o0 == o_begin && t < 2) {
      float v;
      if (o_begin == 0) {
        v = a.prev_in[2 * b + t];
      } else {
        const float *src = reinterpret_cast<const float *>(xiq) + t;
        float acc = 0.0f;  // output o0-1: newest sample is tile sample -D
#pragma unroll 1
        for (int n = 0; n < T; ++n) acc = xmac(acc, taps.h[n], src[2 * Geom::pos(-D - n)]);
        v = xmul(acc, 0.0078125f);
      }
      (t == 0 ? edge_i : edge_q)[0] = v;
    }
    // ---- R outputs per thread, I and Q in the two lanes of a pair ----
    f32x2_t acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0ull;
    fir_window_pairs<T, D, R, HALO, Geom>(xiq + Geom::thread_base(t), taps, acc, one);
    float ai[R], aq[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {  // exact power-of-two scaling: (u8-128)/128
      unpack2(acc[r], ai[r], aq[r]);
      ai[r] = xmul(ai[r], 0.0078125f);
      aq[r] = xmul(aq[r], 0.0078125f);
    }
    edge_i[t + 1] = ai[R - 1];
    edge_q[t + 1] = aq[R - 1];
    __syncthreads();
    float pi = edge_i[t], pq = edge_q[t];
    const int o = o0 + t * R;
    float *drow = a.demod + (size_t)b * a.demod_stride + a.demod_off;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (o + r < o_end) {
        drow[o + r] = fm_demod_one(ai[r], aq[r], pi, pq);
        if (a.i_filt) {
          a.i_filt[(size_t)b * a.tap_stride + o + r] = ai[r];
          a.q_filt[(size_t)b * a.tap_stride + o + r] = aq[r];
        }
        if (o + r == a.n_if - 1) {
          a.prev_out[2 * b] = ai[r];
          a.prev_out[2 * b + 1] = aq[r];
        }
      }
      pi = ai[r];
      pq = aq[r];
    }
  }
}

// Generic-taps fallback of K1, part 1: one thread per output, taps in shared
// memory, run-time loops.  Writes I_filt/Q_filt (slot 0 of each row is prev).
struct RfGenericArgs {
  RfArgs a;
  const float *h;  // [T]
  int T, D;
  float *iq_filt;  // [B][2][1 + n_if]: I row then Q row, slot 0 = previous output
};

static __global__ void k_rf_generic(const RfGenericArgs g) {
  extern __shared__ float sh[];
  for (int n = threadIdx.x; n < g.T; n += blockDim.x) sh[n] = g.h[n];
  __syncthreads();
  const RfArgs &a = g.a;
  const int b = blockIdx.y;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= a.n_if) return;
  const uint8_t *row = a.iq + (size_t)b * a.iq_stride;
  const uint8_t *hrow = a.hist + (size_t)b * 2 * a.rf_hist_len;
  float ai = 0.0f, aq = 0.0f;
  const long long m = (long long)o * g.D;
  for (int n = 0; n < g.T; ++n) {
    float fi, fq;
    rf_fetch_pair<0, 0>(a, row, hrow, m - n, fi, fq);
    // the generic path applies /128 per sample, exactly as iofunc.cpp:133 does
    ai = xmac(ai, sh[n], xmul(fi, 0.0078125f));
    aq = xmac(aq, sh[n], xmul(fq, 0.0078125f));
  }
  float *irow = g.iq_filt + ((size_t)b * 2) * (a.n_if + 1);
  float *qrow = irow + (a.n_if + 1);
  irow[1 + o] = ai;
  qrow[1 + o] = aq;
  if (o == 0) {
    irow[0] = a.prev_in[2 * b];
    qrow[0] = a.prev_in[2 * b + 1];
  }
  if (o == a.n_if - 1) {
    a.prev_out[2 * b] = ai;
    a.prev_out[2 * b + 1] = aq;
  }
  if (a.i_filt) {
    a.i_filt[(size_t)b * a.tap_stride + o] = ai;
    a.q_filt[(size_t)b * a.tap_stride + o] = aq;
  }
}

// Part 2 (also the stand-alone fmDemod operator, filter.cpp:248-266): element-wise
// discriminator.  The one-sample state comes from the neighbouring lane by warp
// shuffle; only lane 0 of each warp re-reads its predecessor from memory.
static __global__ void k_fm_demod(const float *__restrict__ iq_filt, int n, float *__restrict__ out,
                           size_t out_stride, int out_off) {
  const int b = blockIdx.y;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  const float *irow = iq_filt + ((size_t)b * 2) * (n + 1);
  const float *qrow = irow + (n + 1);
  const bool live = o < n;
  float i = live ? irow[1 + o] : 0.0f, q = live ? qrow[1 + o] : 0.0f;
  float pi = __shfl_up_sync(0xffffffffu, i, 1), pq = __shfl_up_sync(0xffffffffu, q, 1);
  if ((threadIdx.x & 31) == 0 && live) {
    pi = irow[o];
    pq = qrow[o];
  }
  if (live) out[(size_t)b * out_stride + out_off + o] = fm_demod_one(i, q, pi, pq);
}

// ---------------------------------------------------------------------------
// K3/K6 (modes 0/1): audio low-pass + decimate + PCM.  Mono: one input (demod).
// Stereo: the mono path reads demod delayed by `delay` samples (allPass,
// filter.cpp:14-29, is a pure delay), the stereo path reads the mixer
// stereo_filt*nco*2 formed on the fly (project.cpp:246-248); outputs L = s+m,
// R = m-s (project.cpp:277-280) as interleaved int16 (project.cpp:294-301).
// ---------------------------------------------------------------------------
struct AudioArgs {
  const float *demod;  // [B][demod_stride], sample 0 at demod_off
  size_t demod_stride;
  int demod_off;
  int delay;           // 0 mono; (stereo_taps-1)/2 stereo
  const float *stf;    // stereo only: [B][stf_stride], sample 0 at hist_off
  const float *nco;    // stereo only: same geometry as stf
  size_t stf_stride, nco_stride;
  int hist_off;
  int16_t *pcm;        // [B][pcm_stride]
  size_t pcm_stride;
  float *audio_filt, *stereo_final;  // optional [B][tap_stride]
  size_t tap_stride;
  int n_out;           // audio samples per capture in this call
  int outs_per_seg;
};

template <int T, int D, int R, int NT>
struct AudioCfg {
  static constexpr int HALO = round_up(T - 1, 4);
  static constexpr int TILE_OUT = NT * R;
  static constexpr int TILE_IN = TILE_OUT * D;
  using Geom = RowGeom<D, R, NT, HALO>;
  static constexpr int ROW = round_up(Geom::FLOATS, 4);
};

template <int T, int D, int R, int NT, bool STEREO, bool FMA = false>
__global__ void __launch_bounds__(NT)
k_audio_fir(const AudioArgs a, const __grid_constant__ TapArray<taps_window(T)> taps) {
  using Cfg = AudioCfg<T, D, R, NT>;
  using Geom = typename Cfg::Geom;
  constexpr int HALO = Cfg::HALO;
  extern __shared__ __align__(16) float smem[];
  float *xm = smem;             // mono-path input
  float *xs = smem + Cfg::ROW;  // stereo-path input (mixer)
  const int t = threadIdx.x;
  const int b = blockIdx.y;
  const int o_begin = blockIdx.x * a.outs_per_seg;
  const int o_end = min(o_begin + a.outs_per_seg, a.n_out);
  const float *drow = a.demod + (size_t)b * a.demod_stride + a.demod_off - a.delay;
  const float *srow = STEREO ? a.stf + (size_t)b * a.stf_stride + a.hist_off : nullptr;
  const float *nrow = STEREO ? a.nco + (size_t)b * a.nco_stride + a.hist_off : nullptr;
  const long long n_in = (long long)a.n_out * D;

  for (int o0 = o_begin; o0 < o_end; o0 += Cfg::TILE_OUT) {
    __syncthreads();
    const long long s0 = (long long)o0 * D - HALO;
    // mono, unpadded rows, 16-byte aligned source, tile inside the call: asynchronous 16-byte
    // copies (a third of this kernel's instructions were the scalar fill below)
    const bool bulk = !STEREO && Geom::PAD == 0 && (a.demod_stride & 3) == 0 &&
                      (((long long)a.demod_off - a.delay + s0) & 3) == 0 && s0 + HALO + Cfg::TILE_IN <= n_in;
    if (bulk) {
      static_assert((HALO + Cfg::TILE_IN) % 4 == 0, "tile is a whole number of float4");
      const float *src = drow + s0;
      for (int q = 4 * t; q < HALO + Cfg::TILE_IN; q += 4 * NT)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(xm + Geom::pos(q - HALO))),
                     "l"(src + q)
                     : "memory");
      asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    } else {
      for (int q = t; q < HALO + Cfg::TILE_IN; q += NT) {
        const long long i = s0 + q;
        const bool ok = i < n_in;  // history prefix makes negative indices valid
        const int pq = Geom::pos(q - HALO);
        xm[pq] = ok ? drow[i] : 0.0f;
        if (STEREO) xs[pq] = ok ? xmul(xmul(srow[i], nrow[i]), 2.0f) : 0.0f;
      }
    }
    __syncthreads();
    float am[R], as[R];
#pragma unroll
    for (int r = 0; r < R; ++r) am[r] = as[r] = 0.0f;
    fir_window<T, D, R, HALO, FMA>(xm + Geom::thread_base(t), taps, am);
    if (STEREO) fir_window<T, D, R, HALO, FMA>(xs + Geom::thread_base(t), taps, as);
    const int o = o0 + t * R;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (o + r < o_end) {
        if (STEREO) {
          const float L = xadd(as[r], am[r]), Rr = xsub(am[r], as[r]);
          int16_t *p = a.pcm + (size_t)b * a.pcm_stride + 2 * (size_t)(o + r);
          p[0] = pcm16(L);
          p[1] = pcm16(Rr);
          if (a.stereo_final) a.stereo_final[(size_t)b * a.tap_stride + o + r] = as[r];
        } else {
          a.pcm[(size_t)b * a.pcm_stride + o + r] = pcm16(am[r]);
        }
        if (a.audio_filt) a.audio_filt[(size_t)b * a.tap_stride + o + r] = am[r];
      }
    }
  }
}

// Stereo audio stage, packed form (EXACT variant): the mono-path and the stereo-path inputs of an
// instant sit side by side in shared memory and share every tap, exactly like I and Q in the front
// end: one FMUL2 + FFMA2 per tap and output for the two filters (project.cpp:219 and :257).
template <int T, int D, int R, int NT>
struct AudioPairCfg {
  static constexpr int HALO = round_up(T - 1, 8);
  static constexpr int TILE_OUT = NT * R;
  static constexpr int TILE_IN = TILE_OUT * D;
  using Geom = RowGeomIQ<D, R, NT, HALO>;
  static constexpr size_t SMEM = (size_t)Geom::PAIRS * sizeof(float2);
};

template <int T, int D, int R, int NT>
__global__ void __launch_bounds__(NT)
k_audio_fir_pair(const AudioArgs a, const __grid_constant__ TapArray<taps_window(T)> taps, float one_rt) {
  using Cfg = AudioPairCfg<T, D, R, NT>;
  using Geom = typename Cfg::Geom;
  constexpr int HALO = Cfg::HALO;
  extern __shared__ __align__(16) float2 xms[];   // (mono-path input, stereo-path input = mixer output)
  const int t = threadIdx.x;
  const int b = blockIdx.y;
  const int o_begin = blockIdx.x * a.outs_per_seg;
  const int o_end = min(o_begin + a.outs_per_seg, a.n_out);
  const float *drow = a.demod + (size_t)b * a.demod_stride + a.demod_off - a.delay;
  const float *srow = a.stf + (size_t)b * a.stf_stride + a.hist_off;
  const float *nrow = a.nco + (size_t)b * a.nco_stride + a.hist_off;
  const long long n_in = (long long)a.n_out * D;
  const f32x2_t one = pack2(one_rt, one_rt);

  for (int o0 = o_begin; o0 < o_end; o0 += Cfg::TILE_OUT) {
    __syncthreads();
    const long long s0 = (long long)o0 * D - HALO;
    for (int q = t; q < HALO + Cfg::TILE_IN; q += NT) {
      const long long i = s0 + q;
      // the history prefixes make negative indices valid down to -(T-1); HALO is rounded up past that
      // for alignment and those first slots are never read by the filter
      const bool ok = i < n_in && q >= HALO - (T - 1);
      xms[Geom::pos(q - HALO)] = ok ? make_float2(drow[i], xmul(xmul(srow[i], nrow[i]), 2.0f)) : make_float2(0.0f, 0.0f);
    }
    __syncthreads();
    f32x2_t acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0ull;
    fir_window_pairs<T, D, R, HALO, Geom>(xms + Geom::thread_base(t), taps, acc, one);
    const int o = o0 + t * R;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (o + r < o_end) {
        float am, as;
        unpack2(acc[r], am, as);
        const float L = xadd(as, am), Rr = xsub(am, as);
        int16_t *p = a.pcm + (size_t)b * a.pcm_stride + 2 * (size_t)(o + r);
        p[0] = pcm16(L);
        p[1] = pcm16(Rr);
        if (a.stereo_final) a.stereo_final[(size_t)b * a.tap_stride + o + r] = as;
        if (a.audio_filt) a.audio_filt[(size_t)b * a.tap_stride + o + r] = am;
      }
    }
  }
}

// Generic-taps form (modes 0/1 with unusual tap counts; also convolveBlockFIR /
// convolveBlockFastFIR as stand-alone operators): one thread per output.
struct FirGenericArgs {
  const float *x;  // [B][x_stride], sample 0 at x_off (history prefix before it)
  size_t x_stride;
  int x_off;
  const float *h;
  int T, D;
  float *y;  // [B][y_stride], output 0 at y_off
  size_t y_stride;
  int y_off;
  int n_out;
};

static __global__ void k_fir_generic(const FirGenericArgs g) {
  extern __shared__ float sh[];
  for (int n = threadIdx.x; n < g.T; n += blockDim.x) sh[n] = g.h[n];
  __syncthreads();
  const int b = blockIdx.y;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= g.n_out) return;
  const float *x = g.x + (size_t)b * g.x_stride + g.x_off + (long long)o * g.D;
  float acc = 0.0f;
  for (int n = 0; n < g.T; ++n) acc = xmac(acc, sh[n], x[-n]);
  g.y[(size_t)b * g.y_stride + g.y_off + o] = acc;
}

// ---------------------------------------------------------------------------
// K6r (modes 2/3): polyphase up/down resampler (convolveBlockResampleFIR,
// filter.cpp:191-223).  Output j sits at upsampled position m = j*D: phase
// p = m % U, newest input i0 = m / U, y = sum_k hp[p][k] * x[i0-k], then
// y += y*U (the reference's (1+U) gain, filter.cpp:213).  hp is the phase-major
// rearrangement hp[p][k] = h[p + k*U]; the zero-stuffed state of the reference
// collapses to the TA-1 sample history prefix.
// ---------------------------------------------------------------------------
struct ResampleArgs {
  AudioArgs a;
  const float *hp;  // [U][TA]
  int U, D, TA;
};

// K6r, throughput form: one LANE per capture.  Consecutive outputs of one capture use
// different polyphase rows and input offsets that are not multiples of anything useful,
// so with lanes along time every shared-memory access conflicts.  With 32 captures across
// the lanes instead, the tap is the same for the whole warp (one broadcast LDS.128 brings
// four taps) and the 32 inputs x[c][i] sit in 32 different banks of a transposed tile.
// CTA = 32 captures x RS_J outputs; warp w computes outputs w, w+NW, ... of the tile.
template <bool FMA>
__device__ __forceinline__ float rs_mac(float acc, float h, float x) {
  return FMA ? __fmaf_rn(h, x, acc) : xmac(acc, h, x);
}
constexpr int RS_J = 32;    // outputs per tile
constexpr int RS_NW = 8;    // warps per CTA
constexpr int RS_PITCH = 33;

template <bool STEREO, bool FMA = false>
static __global__ void __launch_bounds__(RS_NW * 32)
k_audio_resample_v2(const ResampleArgs g, int batch, int rows_cap) {
  const AudioArgs &a = g.a;
  extern __shared__ __align__(16) float smem[];
  const int TA4 = (g.TA + 3) & ~3;
  float *hs = smem;                                  // [RS_J][TA4] taps of the tile's phases
  float *xs = hs + RS_J * TA4;                       // [rows_cap][33] mono-path input, transposed
  float *xs2 = xs + (size_t)rows_cap * RS_PITCH;     // stereo-path input (mixer), transposed
  int16_t *ps = reinterpret_cast<int16_t *>(xs + (size_t)rows_cap * RS_PITCH * (STEREO ? 2 : 1));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j0 = blockIdx.x * RS_J;
  const int jn = min(RS_J, a.n_out - j0);
  const int c0 = blockIdx.y * 32;
  // All positions fit in 32 bits: the pipeline caps n_if * U below 2^31.
  const unsigned U = (unsigned)g.U, D = (unsigned)g.D;
  const unsigned m0 = (unsigned)j0 * D;
  const int i_lo = (int)(m0 / U) - (g.TA - 1);                          // oldest input of the tile
  const int i_hi = (int)(((unsigned)(j0 + jn - 1) * D) / U);            // newest input of the tile
  const int rows = i_hi - i_lo + 1;
  __shared__ int s_phase[RS_J], s_top[RS_J];
  if (threadIdx.x < RS_J) {
    const unsigned m = (unsigned)(j0 + threadIdx.x) * D;
    const unsigned q = m / U;
    s_phase[threadIdx.x] = (int)(m - q * U);
    s_top[threadIdx.x] = (int)q - i_lo;   // row of the newest input of this output
  }
  __syncthreads();
  // ---- taps of the jn phases, zero padded to TA4 ----
  for (int jj = warp; jj < jn; jj += RS_NW) {
    const float *src = g.hp + (size_t)s_phase[jj] * g.TA;
    for (int k = lane; k < TA4; k += 32) hs[jj * TA4 + k] = (k < g.TA) ? __ldg(src + k) : 0.0f;
  }
  // ---- transposed input tiles: each warp streams whole rows of its captures ----
  for (int c = warp; c < 32; c += RS_NW) {
    const int ch = c0 + c;
    if (ch < batch) {
      const float *drow = a.demod + (size_t)ch * a.demod_stride + a.demod_off - a.delay + i_lo;
      const float *srow = STEREO ? a.stf + (size_t)ch * a.stf_stride + a.hist_off + i_lo : nullptr;
      const float *nrow = STEREO ? a.nco + (size_t)ch * a.nco_stride + a.hist_off + i_lo : nullptr;
      for (int i = lane; i < rows; i += 32) {
        xs[i * RS_PITCH + c] = drow[i];
        if (STEREO) xs2[i * RS_PITCH + c] = xmul(xmul(srow[i], nrow[i]), 2.0f);
      }
    } else {
      for (int i = lane; i < rows; i += 32) {
        xs[i * RS_PITCH + c] = 0.0f;
        if (STEREO) xs2[i * RS_PITCH + c] = 0.0f;
      }
    }
  }
  __syncthreads();
  const float fu = (float)g.U;
  for (int jj = warp; jj < jn; jj += RS_NW) {
    const int top = s_top[jj];
    const float *h = hs + jj * TA4;
    const float *x = xs + top * RS_PITCH + lane;
    const float *x2 = xs2 + top * RS_PITCH + lane;
    float am = 0.0f, as = 0.0f;
    int k = 0;
    for (; k + 4 <= g.TA; k += 4) {
      const float4 hv = *reinterpret_cast<const float4 *>(h + k);
      am = rs_mac<FMA>(am, hv.x, x[-(k + 0) * RS_PITCH]);
      am = rs_mac<FMA>(am, hv.y, x[-(k + 1) * RS_PITCH]);
      am = rs_mac<FMA>(am, hv.z, x[-(k + 2) * RS_PITCH]);
      am = rs_mac<FMA>(am, hv.w, x[-(k + 3) * RS_PITCH]);
      if (STEREO) {
        as = rs_mac<FMA>(as, hv.x, x2[-(k + 0) * RS_PITCH]);
        as = rs_mac<FMA>(as, hv.y, x2[-(k + 1) * RS_PITCH]);
        as = rs_mac<FMA>(as, hv.z, x2[-(k + 2) * RS_PITCH]);
        as = rs_mac<FMA>(as, hv.w, x2[-(k + 3) * RS_PITCH]);
      }
    }
    for (; k < g.TA; ++k) {
      am = rs_mac<FMA>(am, h[k], x[-k * RS_PITCH]);
      if (STEREO) as = rs_mac<FMA>(as, h[k], x2[-k * RS_PITCH]);
    }
    am = xadd(am, xmul(am, fu));  // filter.cpp:213
    if (STEREO) as = xadd(as, xmul(as, fu));
    const int ch = c0 + lane;
    if (STEREO) {
      ps[(lane * RS_J + jj) * 2] = pcm16(xadd(as, am));
      ps[(lane * RS_J + jj) * 2 + 1] = pcm16(xsub(am, as));
    } else {
      ps[lane * RS_J + jj] = pcm16(am);
    }
    if (ch < batch) {
      if (a.audio_filt) a.audio_filt[(size_t)ch * a.tap_stride + j0 + jj] = am;
      if (STEREO && a.stereo_final) a.stereo_final[(size_t)ch * a.tap_stride + j0 + jj] = as;
    }
  }
  __syncthreads();
  // ---- PCM rows out: one capture per warp pass, contiguous int16 along time ----
  constexpr int PER = STEREO ? 2 : 1;
  for (int c = warp; c < 32; c += RS_NW) {
    const int ch = c0 + c;
    if (ch >= batch) continue;
    int16_t *dst = a.pcm + (size_t)ch * a.pcm_stride + (size_t)j0 * PER;
    for (int q = lane; q < jn * PER; q += 32) dst[q] = ps[c * RS_J * PER + q];
  }
}

// Pair form (stereo, 101 taps per phase): the generic form above sits at 79 % of the shared-memory
// wavefront peak (profiles/r1f) because every multiply-add reads its own input word.  Two CONSECUTIVE outputs use input windows that are
// only DU0 or DU0+1 = floor(D/U) (+1) rows apart, so a warp that computes them together loads each
// input row once and feeds both accumulators: 0.79 instead of 1.25 shared-memory wavefronts per
// multiply-add.  The row walk is unrolled for both possible offsets (a warp-uniform branch picks
// one); every accumulator still sees its taps in ascending order.
template <bool STEREO, int TA, int DELTA, bool FMA>
__device__ __forceinline__ void rs_pair(const float *__restrict__ ha, const float *__restrict__ hb,
                                        const float *__restrict__ xtop, const float *__restrict__ ytop,
                                        float &ma, float &mb, float &sa, float &sb) {
  // xtop/ytop point at the newest row of output b (this lane's column); output a starts DELTA rows down
  float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
#pragma unroll
  for (int kb = 0; kb < TA + DELTA; ++kb) {
    const int ka = kb - DELTA;
    if (kb < TA && (kb & 3) == 0) vb = *reinterpret_cast<const float4 *>(hb + kb);
    if (ka >= 0 && ka < TA && (ka & 3) == 0) va = *reinterpret_cast<const float4 *>(ha + ka);
    const float x = xtop[-kb * RS_PITCH];
    const float y = STEREO ? ytop[-kb * RS_PITCH] : 0.0f;
    if (kb < TA) {
      const float t = (kb & 3) == 0 ? vb.x : (kb & 3) == 1 ? vb.y : (kb & 3) == 2 ? vb.z : vb.w;
      mb = rs_mac<FMA>(mb, t, x);
      if (STEREO) sb = rs_mac<FMA>(sb, t, y);
    }
    if (ka >= 0 && ka < TA) {
      const float t = (ka & 3) == 0 ? va.x : (ka & 3) == 1 ? va.y : (ka & 3) == 2 ? va.z : va.w;
      ma = rs_mac<FMA>(ma, t, x);
      if (STEREO) sa = rs_mac<FMA>(sa, t, y);
    }
  }
}

template <bool STEREO, int TA, int DU0, bool FMA = false>
static __global__ void __launch_bounds__(RS_NW * 32)
k_audio_resample_v4(const ResampleArgs g, int batch, int rows_cap) {
  const AudioArgs &a = g.a;
  constexpr int TA4 = (TA + 3) & ~3;
  constexpr int PER = STEREO ? 2 : 1;
  constexpr int PSP = RS_J * PER + 2;                // padded PCM row (int16) -> fewer bank conflicts
  extern __shared__ __align__(16) float smem[];
  float *hs = smem;                                  // [RS_J][TA4]
  float *xs = hs + RS_J * TA4;                       // [rows_cap][33]
  float *xs2 = xs + (size_t)rows_cap * RS_PITCH;
  int16_t *ps = reinterpret_cast<int16_t *>(xs + (size_t)rows_cap * RS_PITCH * PER);
  __shared__ int s_phase[RS_J], s_top[RS_J];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j0 = blockIdx.x * RS_J;
  const int jn = min(RS_J, a.n_out - j0);
  const int c0 = blockIdx.y * 32;
  const unsigned U = (unsigned)g.U, D = (unsigned)g.D;
  const int i_lo = (int)(((unsigned)j0 * D) / U) - (TA - 1);
  const int i_hi = (int)(((unsigned)(j0 + jn - 1) * D) / U);
  const int rows = i_hi - i_lo + 1;
  if (threadIdx.x < RS_J) {
    const unsigned m = (unsigned)(j0 + threadIdx.x) * D;
    const unsigned q = m / U;
    s_phase[threadIdx.x] = (int)(m - q * U);
    s_top[threadIdx.x] = (int)q - i_lo;   // also for outputs past n_out: rows_cap covers a full tile
  }
  // ---- transposed input tile: warp w streams the rows of captures w, w+8, ... ----
#pragma unroll
  for (int cc = 0; cc < 32 / RS_NW; ++cc) {
    const int c = warp + cc * RS_NW;
    const int ch = min(c0 + c, batch - 1);           // lanes past the batch re-read the last capture
    const float *drow = a.demod + (size_t)ch * a.demod_stride + (a.demod_off - a.delay + i_lo);
    float *dst = xs + c + lane * RS_PITCH;
    if (STEREO) {
      const float *srow = a.stf + (size_t)ch * a.stf_stride + (a.hist_off + i_lo);
      const float *nrow = a.nco + (size_t)ch * a.nco_stride + (a.hist_off + i_lo);
      float *dst2 = xs2 + c + lane * RS_PITCH;
      for (int i = lane; i < rows; i += 32, dst += 32 * RS_PITCH, dst2 += 32 * RS_PITCH) {
        *dst = drow[i];
        *dst2 = xmul(xmul(srow[i], nrow[i]), 2.0f);
      }
    } else {
      int i = lane;
      for (; i + 96 < rows; i += 128, dst += 128 * RS_PITCH) {  // four loads in flight
        const float v0 = drow[i], v1 = drow[i + 32], v2 = drow[i + 64], v3 = drow[i + 96];
        dst[0] = v0; dst[32 * RS_PITCH] = v1; dst[64 * RS_PITCH] = v2; dst[96 * RS_PITCH] = v3;
      }
      for (; i < rows; i += 32, dst += 32 * RS_PITCH) *dst = drow[i];
    }
  }
  __syncthreads();   // s_phase visible
  // ---- taps of the tile's phases (zero padded to TA4) ----
  for (int jj = warp; jj < RS_J; jj += RS_NW) {
    const float *src = g.hp + (size_t)s_phase[jj] * TA;
#pragma unroll
    for (int k0 = 0; k0 < TA4; k0 += 32) {
      const int k = k0 + lane;
      if (k < TA4) hs[jj * TA4 + k] = (k < TA) ? __ldg(src + k) : 0.0f;
    }
  }
  __syncthreads();
  const float fu = (float)g.U;
  // ---- each warp: consecutive outputs (jj, jj+1) together, 32 captures across the lanes ----
  for (int jj = 2 * warp; jj < RS_J; jj += 2 * RS_NW) {
    const int jb = jj + 1;
    const float *ha = hs + jj * TA4, *hb = hs + jb * TA4;
    const int top_b = s_top[jb], delta = top_b - s_top[jj];
    const float *xtop = xs + top_b * RS_PITCH + lane, *ytop = xs2 + top_b * RS_PITCH + lane;
    float ma = 0.0f, mb = 0.0f, sa = 0.0f, sb = 0.0f;
    if (delta == DU0) rs_pair<STEREO, TA, DU0, FMA>(ha, hb, xtop, ytop, ma, mb, sa, sb);
    else rs_pair<STEREO, TA, DU0 + 1, FMA>(ha, hb, xtop, ytop, ma, mb, sa, sb);
    ma = xadd(ma, xmul(ma, fu));  // filter.cpp:213
    mb = xadd(mb, xmul(mb, fu));
    if (STEREO) {
      sa = xadd(sa, xmul(sa, fu));
      sb = xadd(sb, xmul(sb, fu));
      ps[lane * PSP + 2 * jj] = pcm16(xadd(sa, ma));
      ps[lane * PSP + 2 * jj + 1] = pcm16(xsub(ma, sa));
      ps[lane * PSP + 2 * jb] = pcm16(xadd(sb, mb));
      ps[lane * PSP + 2 * jb + 1] = pcm16(xsub(mb, sb));
    } else {
      ps[lane * PSP + jj] = pcm16(ma);
      ps[lane * PSP + jb] = pcm16(mb);
    }
    const int ch = c0 + lane;
    if (ch < batch && a.audio_filt) {
      if (jj < jn) a.audio_filt[(size_t)ch * a.tap_stride + j0 + jj] = ma;
      if (jb < jn) a.audio_filt[(size_t)ch * a.tap_stride + j0 + jb] = mb;
      if (STEREO && a.stereo_final) {
        if (jj < jn) a.stereo_final[(size_t)ch * a.tap_stride + j0 + jj] = sa;
        if (jb < jn) a.stereo_final[(size_t)ch * a.tap_stride + j0 + jb] = sb;
      }
    }
  }
  __syncthreads();
  // ---- PCM rows out: one capture per warp pass, contiguous int16 along time ----
  for (int c = warp; c < 32; c += RS_NW) {
    const int ch = c0 + c;
    if (ch >= batch) continue;
    int16_t *dst = a.pcm + (size_t)ch * a.pcm_stride + (size_t)j0 * PER;
    for (int q = lane; q < jn * PER; q += 32) dst[q] = ps[c * PSP + q];
  }
}

// Quad form (mono default).  profiles/r1h: the pair form spends 40 % of its time filling the
// transposed tile (load -> shared store round trips) and its multiply-add loop is bound by
// shared-memory wavefronts (0.78 per multiply-add).  Here
//  * the input tile stays in its natural [capture][time] layout with an ODD row pitch, so it is
//    filled by 4-byte asynchronous copies (no registers, no transposition) and lane c still reads
//    bank (c*pitch + t) % 32: conflict free;
//  * a warp computes FOUR consecutive outputs for 64 captures (two per lane): one input word
//    feeds four accumulators, and the four outputs' taps arrive as ONE broadcast float4 from a
//    per-phase table that the host lays out in the warp's walk order (`tq`, below): 3 wavefronts
//    per 8 multiply-adds.
// Every accumulator still sees its own taps in ascending order (the table only inserts zero
// taps before/after them, and acc + 0*x == acc), so the result stays bit-identical to
// filter.cpp:191-223; FMA=true (SDR_VARIANT_FAST only) contracts the multiply-add.
//
// tq[phi0][kb][o], phi0 = phase of the quad's first output, o = 0..3:
//   tq = hp[phase_o][kb - d_o] (0 outside 0..TA-1),  d_o = top_3 - top_o,  input row = top_3 - kb.
constexpr int RQ_J = 32;        // outputs per tile (8 warps x 4)
struct ResampleQuadArgs {
  AudioArgs a;
  const float *tq;  // [U][KB][4]
  int U, D, TA, KB;
  int n_in;         // valid input samples of this call (rows at or past it read as zero)
};

template <bool FMA>
__device__ __forceinline__ float rq_mac(float acc, float h, float x) {
  return FMA ? __fmaf_rn(h, x, acc) : xmac(acc, h, x);
}

// Tile rows are 16-byte aligned in shared AND global memory (the tile starts at a multiple of
// four samples of the capture's buffer), so they are filled by 16-byte asynchronous copies and a
// lane reads FOUR consecutive input samples with one LDS.128; pitch/4 is odd, which keeps a
// quarter-warp's LDS.128 on eight distinct 16-byte bank groups.  A quad whose newest row is not
// the last of an aligned group walks `r` rows earlier; its tap table is staged r rows lower
// (rows are float4, so the shift keeps the copies aligned) with zero rows around it.
template <bool FMA, int NC>
static __global__ void __launch_bounds__(256)
k_audio_resample_v5(const ResampleQuadArgs g, int batch, int pitch) {
  const AudioArgs &a = g.a;
  constexpr int CAPS = 32 * NC;                       // captures per tile (NC per lane)
  constexpr int PSP = RQ_J + 2;                       // padded PCM row (int16)
  const int KBP = g.KB + 4;                           // table rows incl. the alignment shift
  extern __shared__ __align__(16) float smem[];
  float *tqs = smem;                                  // [8 warps][KBP][4]
  float *xs = tqs + 8 * KBP * 4;                      // [CAPS][pitch]
  int16_t *ps = reinterpret_cast<int16_t *>(xs + (size_t)CAPS * pitch);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j0 = blockIdx.x * RQ_J;
  const int jn = min(RQ_J, a.n_out - j0);
  const int c0 = blockIdx.y * CAPS;
  const unsigned U = (unsigned)g.U, D = (unsigned)g.D;
  // rows [i_lo, i_hi]: from the oldest row any quad of the tile can touch (KBP-1 below its
  // aligned newest row), moved down to a multiple of four samples of the buffer, to the end of
  // the aligned group holding the newest row of the last quad with a valid output
  const int off = a.demod_off - a.delay;              // buffer index of input sample 0
  int i_lo = (int)(((unsigned)j0 * D) / U) - (KBP - 1);
  i_lo -= (((off + i_lo) % 4) + 4) % 4;
  const int jlast = j0 + ((jn + 3) & ~3) - 1;
  const int i_hi = ((int)(((unsigned)jlast * D) / U) - i_lo) | 3;   // relative to i_lo, end of its group
  const int rows = i_hi + 1;                          // multiple of 4
  const int first = off + i_lo;                       // buffer index of tile row 0 (multiple of 4)
  const bool interior = first >= 0 && i_lo + i_hi < g.n_in && (a.demod_stride & 3) == 0;
  const uint32_t xs_s = (uint32_t)__cvta_generic_to_shared(xs);
  for (int c = warp; c < CAPS; c += 8) {
    const int ch = min(c0 + c, batch - 1);            // lanes past the batch re-read the last capture
    const float *src = a.demod + (size_t)ch * a.demod_stride + first;
    const uint32_t dst = xs_s + (uint32_t)(c * pitch) * 4u;
    if (interior) {
      for (int i = 4 * lane; i < rows; i += 128)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 4u * i), "l"(src + i) : "memory");
    } else {
      for (int i = lane; i < rows; i += 32)
        xs[(size_t)c * pitch + i] = (first + i >= 0 && i_lo + i < g.n_in) ? src[i] : 0.0f;
    }
  }
  const int jq = j0 + 4 * warp;
  const bool active = jq < a.n_out;
  int gtop = 0;
  if (active) {
    const int top3 = (int)(((unsigned)(jq + 3) * D) / U) - i_lo;
    gtop = top3 | 3;                                  // newest row of the aligned group
    const int r = gtop - top3;                        // rows walked before the quad's own newest row
    const unsigned phi0 = ((unsigned)jq * D) % U;
    const float4 *src = reinterpret_cast<const float4 *>(g.tq) + (size_t)phi0 * g.KB;
    float4 *dst = reinterpret_cast<float4 *>(tqs) + warp * KBP;
    for (int i = lane; i < KBP; i += 32) {
      const int k = i - r;
      if (k >= 0 && k < g.KB)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + i)),
                     "l"(src + k) : "memory");
      else
        dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (active) {
    const float *x0 = xs + (size_t)lane * pitch + gtop - 3;   // aligned group holding the newest row
    const float4 *t = reinterpret_cast<const float4 *>(tqs) + warp * KBP;
    float acc[4][NC] = {};
#pragma unroll 2
    for (int kb = 0; kb < KBP; kb += 4) {
      float4 xv[NC];
#pragma unroll
      for (int cc = 0; cc < NC; ++cc) xv[cc] = *reinterpret_cast<const float4 *>(x0 + (size_t)32 * cc * pitch - kb);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 h = t[kb + u];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
          const float x = u == 0 ? xv[cc].w : u == 1 ? xv[cc].z : u == 2 ? xv[cc].y : xv[cc].x;
          acc[0][cc] = rq_mac<FMA>(acc[0][cc], h.x, x);
          acc[1][cc] = rq_mac<FMA>(acc[1][cc], h.y, x);
          acc[2][cc] = rq_mac<FMA>(acc[2][cc], h.z, x);
          acc[3][cc] = rq_mac<FMA>(acc[3][cc], h.w, x);
        }
      }
    }
    const float fu = (float)g.U;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
#pragma unroll
      for (int cc = 0; cc < NC; ++cc) {
        const float y = xadd(acc[o][cc], xmul(acc[o][cc], fu));  // filter.cpp:213
        const int c = lane + 32 * cc;
        ps[c * PSP + 4 * warp + o] = pcm16(y);
        if (a.audio_filt && c0 + c < batch && jq + o < a.n_out)
          a.audio_filt[(size_t)(c0 + c) * a.tap_stride + jq + o] = y;
      }
    }
  }
  __syncthreads();
  // ---- PCM rows out: two captures per warp pass, 32-bit words along time ----
  const bool words = jn == RQ_J && (a.pcm_stride & 1) == 0 && (reinterpret_cast<uintptr_t>(a.pcm) & 3) == 0;
  if (words) {
    for (int c = 2 * warp + (lane >> 4); c < CAPS; c += 16) {
      const int ch = c0 + c;
      if (ch >= batch) continue;
      uint32_t *dst = reinterpret_cast<uint32_t *>(a.pcm + (size_t)ch * a.pcm_stride + (size_t)j0);
      dst[lane & 15] = *reinterpret_cast<const uint32_t *>(ps + c * PSP + 2 * (lane & 15));
    }
  } else {
    for (int c = warp; c < CAPS; c += 8) {
      const int ch = c0 + c;
      if (ch >= batch) break;
      int16_t *dst = a.pcm + (size_t)ch * a.pcm_stride + (size_t)j0;
      if (lane < jn) dst[lane] = ps[c * PSP + lane];
    }
  }
}

// Stand-alone resampler on float in/out (no PCM), for sdr_fir_resample.
struct ResampleOpArgs {
  const float *x;  // sample 0 at x_off, TA-1 history before it
  int x_off;
  const float *hp;
  int U, D, TA;
  float *y;
  int n_out;
};
static __global__ void k_resample_op(const ResampleOpArgs g) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g.n_out) return;
  const long long m = (long long)j * g.D;
  const int p = (int)(m % g.U);
  const float *x = g.x + g.x_off + m / g.U;
  const float *hp = g.hp + (size_t)p * g.TA;
  float acc = 0.0f;
  for (int k = 0; k < g.TA; ++k) acc = xmac(acc, hp[k], x[-k]);
  g.y[j] = xadd(acc, xmul(acc, (float)g.U));
}

// Mono/stereo PCM epilogue for the generic-taps path (float audio -> int16).
static __global__ void k_pcm_pack(const float *mono, const float *stereo, size_t in_stride, int n,
                           int16_t *pcm, size_t pcm_stride) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float m = mono[(size_t)b * in_stride + j];
  if (stereo) {
    const float s = stereo[(size_t)b * in_stride + j];
    pcm[(size_t)b * pcm_stride + 2 * (size_t)j] = pcm16(xadd(s, m));
    pcm[(size_t)b * pcm_stride + 2 * (size_t)j + 1] = pcm16(xsub(m, s));
  } else {
    pcm[(size_t)b * pcm_stride + j] = pcm16(m);
  }
}

// mixer = stereo_filt * nco * 2 into a history-prefixed row (generic path / taps).
static __global__ void k_mixer(const float *stf, size_t stf_stride, const float *nco, size_t nco_stride,
                        int off, int n_total, float *out, size_t out_stride) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_total) return;
  out[(size_t)b * out_stride + i] =
      xmul(xmul(stf[(size_t)b * stf_stride + off + i], nco[(size_t)b * nco_stride + off + i]), 2.0f);
}

// Row-wise copy used to keep intermediates that the carry would overwrite.
static __global__ void k_copy_rows(const float *src, size_t src_stride, int src_off, float *dst,
                            size_t dst_stride, int n) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[(size_t)b * dst_stride + i] = src[(size_t)b * src_stride + src_off + i];
}

// ---------------------------------------------------------------------------
// K4: the two stereo band-pass filters on one staged input tile
// (convolveBlockFIR x2, project.cpp:202,207; filter.cpp:133-154).
// ---------------------------------------------------------------------------
struct BpfArgs {
  const float *demod;
  size_t demod_stride;
  int demod_off;
  float *stf;  // [B][stf_stride], output 0 at hist_off
  size_t stf_stride;
  int hist_off;
  float *car;  // [B][car_stride]
  size_t car_stride;
  int n_if;
  int outs_per_seg;
  float one;   // 1.0f at run time (xmac2)
};

template <int T, int R, int NT, bool FMA_STEREO = false>
__global__ void __launch_bounds__(NT)
k_bpf_dual(const BpfArgs a, const __grid_constant__ TapArray<taps_groups(T, 1, R)> h_stereo,
           const __grid_constant__ TapArray<taps_groups(T, 1, R)> h_pilot) {
  constexpr int HALO = round_up(taps_groups(T, 1, R) + R, 4);
  constexpr int TILE = NT * R;
  using Geom = RowGeom<1, R, NT, HALO>;
  __shared__ __align__(16) float xs[round_up(Geom::FLOATS, 4)];
  const int t = threadIdx.x;
  const int b = blockIdx.y;
  const int o_begin = blockIdx.x * a.outs_per_seg;
  const int o_end = min(o_begin + a.outs_per_seg, a.n_if);
  const float *drow = a.demod + (size_t)b * a.demod_stride + a.demod_off;
  for (int o0 = o_begin; o0 < o_end; o0 += TILE) {
    __syncthreads();
    for (int q = t; q < HALO + TILE; q += NT) {
      const int i = o0 - HALO + q;
      // the padded taps reach further back than the real filter: those samples only meet
      // zero taps, but they must be finite, so positions before the history read as 0
      xs[Geom::pos(q - HALO)] = (i < a.n_if && i >= -a.demod_off) ? drow[i] : 0.0f;
    }
    __syncthreads();
    float as[R], ap[R];
#pragma unroll
    for (int r = 0; r < R; ++r) as[r] = ap[r] = 0.0f;
    fir_groups2<T, 1, R, FMA_STEREO>(xs + Geom::thread_base(t), h_stereo, h_pilot, as, ap);
    const int o = o0 + t * R;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (o + r < o_end) {
        a.stf[(size_t)b * a.stf_stride + a.hist_off + o + r] = as[r];
        a.car[(size_t)b * a.car_stride + o + r] = ap[r];
      }
    }
  }
}

// Packed form (EXACT variant): see fir_groups2_packed.  Same staging, same bits.
template <int T, int R, int NT>
__global__ void __launch_bounds__(NT)
k_bpf_dual_packed(const BpfArgs a, const __grid_constant__ TapPairs<taps_groups(T, 1, R)> h2) {
  constexpr int HALO = round_up(taps_groups(T, 1, R) + R, 4);
  constexpr int TILE = NT * R;
  using Geom = RowGeom<1, R, NT, HALO>;
  __shared__ __align__(16) float xs[round_up(Geom::FLOATS, 4)];
  const int t = threadIdx.x;
  const int b = blockIdx.y;
  const int o_begin = blockIdx.x * a.outs_per_seg;
  const int o_end = min(o_begin + a.outs_per_seg, a.n_if);
  const float *drow = a.demod + (size_t)b * a.demod_stride + a.demod_off;
  const f32x2_t one = pack2(a.one, a.one);
  for (int o0 = o_begin; o0 < o_end; o0 += TILE) {
    __syncthreads();
    for (int q = t; q < HALO + TILE; q += NT) {
      const int i = o0 - HALO + q;
      xs[Geom::pos(q - HALO)] = (i < a.n_if && i >= -a.demod_off) ? drow[i] : 0.0f;
    }
    __syncthreads();
    f32x2_t acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0ull;
    fir_groups2_packed<T, 1, R>(xs + Geom::thread_base(t), h2, acc, one);
    const int o = o0 + t * R;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (o + r < o_end) {
        float vs, vp;
        unpack2(acc[r], vs, vp);
        a.stf[(size_t)b * a.stf_stride + a.hist_off + o + r] = vs;
        a.car[(size_t)b * a.car_stride + o + r] = vp;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// K5: PLL + NCO (fmPLL, filter.cpp:32-80).  The recurrence is inherently
// sequential, so one lane owns one capture and walks its samples in order.
// ncoOut[k+1] = cosf(trigArg*ncoScale + phaseAdjust) (filter.cpp:71) is not fed back, so
// K5a (k_pll) only stores its ARGUMENT and K5b (k_nco_cos) evaluates the cosines for
// all samples in parallel.  In one warp the cosine's ~150 instructions, although off the
// dependency chain, are issued in order between the chain's: 855 cycles per sample with
// it, 590 without (tools/ubench_pll_chain.cu).
// The NCO output is written one slot late (slot k+1 at step k) which is exactly
// the reference's ncoOut vector: ncoOut[0] is the last value of the previous
// block (state[4]) and the mixer reads ncoOut[0..N) (project.cpp:246-248).
// ---------------------------------------------------------------------------
struct PllArgs {
  const float *in;  // [B][in_stride]
  size_t in_stride;
  float *out;  // [B][out_stride]; step k writes out[out_off + k + 1]
  size_t out_stride;
  int out_off;
  float *state;  // [B][8]: integrator, phaseEst, feedbackI, feedbackQ, lastOut, trigOffset
  int n;
  int batch;
  float freq, Fs, ncoScale, phaseAdjust, normBandwidth;
};

constexpr int PLL_MAX_WARPS = 4;   // warps per block at large batches (each warp: 32 captures, own tiles)
static __global__ void __launch_bounds__(32 * PLL_MAX_WARPS) k_pll(const PllArgs a) {
  // lanes past the batch shadow the last capture (same loads, same stores of the same values, in
  // lockstep with its owner): the tile copies below are warp-collective.  A whole warp past the
  // batch has nothing to shadow.
  if ((int)((blockIdx.x * blockDim.x + threadIdx.x) & ~31u) >= a.batch) return;
  const int b = min((int)(blockIdx.x * blockDim.x + threadIdx.x), a.batch - 1);
  // filter.cpp:35-39
  const float Kp = xmul(a.normBandwidth, 2.666f);
  const float Ki = xmul(xmul(a.normBandwidth, a.normBandwidth), 3.555f);
  float *st = a.state + (size_t)b * 8;
  float integrator = st[0], phaseEst = st[1], fbI = st[2], fbQ = st[3], trigOffset = st[5];
  const float *in = a.in + (size_t)b * a.in_stride;
  float *out = a.out + (size_t)b * a.out_stride + a.out_off;
  // filter.cpp:68: 2*PI*(freq/Fs) evaluated in double from the float quotient
  const double w = __dmul_rn(6.283185307179586476925286766559, (double)xdiv(a.freq, a.Fs));
  // Input: the warp copies tiles of 32 captures x 32 samples into shared memory with
  // asynchronous copies (32 coalesced 128-byte rows per tile, no register in between), one
  // tile ahead of the one being consumed, so no global-load latency meets the recurrence.
  __shared__ float tiles[PLL_MAX_WARPS][2][32][33];
  float (*tile)[32][33] = tiles[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int b_first = (blockIdx.x * blockDim.x + threadIdx.x) & ~31;
  auto fetch = [&](int buf, int k0) {
    for (int c = 0; c < 32; ++c) {
      const int bc = min(b_first + c, a.batch - 1);
      const int k = min(k0 + lane, a.n - 1);
      const float *src = a.in + (size_t)bc * a.in_stride + k;
      const unsigned dst = (unsigned)__cvta_generic_to_shared(&tile[buf][c][lane]);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  fetch(0, 0);
  for (int k0 = 0, buf = 0; k0 < a.n; k0 += 32, buf ^= 1) {
    if (k0 + 32 < a.n) {
      fetch(buf ^ 1, k0 + 32);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    const float *xs = tile[buf][lane];
    const int kn = min(32, a.n - k0);
    float x_next = xs[0];
    for (int j = 0; j < kn; ++j) {
      const float x = x_next;
      x_next = xs[min(j + 1, 31)];  // a step ahead: the shared-memory latency stays off the chain
      const float eI = xmul(x, fbI);
      const float eQ = xmul(x, -fbQ);
      float eD;
      if (!atan2f_common(eQ, eI, eD)) eD = atan2f_glibc(eQ, eI);
      integrator = xadd(integrator, xmul(Ki, eD));
      phaseEst = xadd(xadd(phaseEst, xmul(Kp, eD)), integrator);
      trigOffset = xadd(trigOffset, 1.0f);
      const float trigArg =
          __double2float_rn(__dadd_rn(__dmul_rn(w, (double)trigOffset), (double)phaseEst));
      sincosf_glibc_bf(trigArg, fbQ, fbI);
      out[k0 + j + 1] = xadd(xmul(trigArg, a.ncoScale), a.phaseAdjust);  // K5b turns it into ncoOut[k+1]
    }
    __syncwarp();
  }
  st[0] = integrator;
  st[1] = phaseEst;
  st[2] = fbI;
  st[3] = fbQ;
  st[5] = trigOffset;
}

// K5b: ncoOut[k+1] = cosf(argument) in place, four samples per thread; the last one is also the
// next call's ncoOut[0] (state[4], filter.cpp:79).  Throughput-bound, so the branchy form
// (only the reduction path and the polynomial that are needed) is the right one here.
static __global__ void __launch_bounds__(256) k_nco_cos(const PllArgs a) {
  const int k0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int b = blockIdx.y;
  if (k0 >= a.n) return;
  float *slot = a.out + (size_t)b * a.out_stride + a.out_off + k0 + 1;
  float v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = k0 + i < a.n ? slot[i] : 0.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = cosf_glibc(v[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (k0 + i < a.n) {
      slot[i] = v[i];
      if (k0 + i == a.n - 1) a.state[(size_t)b * 8 + 4] = v[i];
    }
}

// ---------------------------------------------------------------------------
// Carry: after a call, move the tails that the next call needs into the history
// prefixes (the reference's "prepare next state" loops, filter.cpp:148-153,
// 183-187, 217-222) and rebuild the raw I/Q history.
// ---------------------------------------------------------------------------
struct CarryArgs {
  // float rows: copy row[off_src .. off_src+len) -> row[0 .. len)
  float *rows[5];      // demod, stereo_filt, nco, and the two fp16 planes of the tensor-core resampler
  size_t strides[5];   // (plane rows are moved as pairs of halfs: offsets and lengths in 32-bit words)
  int src_off[5];
  int len[5];
  // raw I/Q history
  const uint8_t *iq;
  size_t iq_stride;
  uint8_t *hist;
  int rf_hist_len;
  long long n_rf;
  float *prev_dst;
  const float *prev_src;
  int *work_counter;   // optional: work-item counter of the tensor-core front end, re-armed for the next call
};

static __global__ void k_carry(const CarryArgs c) {
  extern __shared__ unsigned char stage[];
  const int b = blockIdx.x;
  const int t = threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (c.work_counter && b == 0 && t == 0) *c.work_counter = 0;
  // 1. float tails.  Source and destination ranges may overlap when the call was
  //    shorter than the history, so go through shared memory.
  float *fstage = reinterpret_cast<float *>(stage);
  for (int r = 0; r < 5; ++r) {
    if (!c.rows[r]) continue;
    float *row = c.rows[r] + (size_t)b * c.strides[r];
    for (int i = t; i < c.len[r]; i += blockDim.x) fstage[i] = row[c.src_off[r] + i];
    __syncthreads();
    for (int i = t; i < c.len[r]; i += blockDim.x) row[i] = fstage[i];
    __syncthreads();
  }
  // 2. raw history: the last HR pairs of (old history ++ this call's input)
  const int HR = c.rf_hist_len;
  uint8_t *hrow = c.hist + (size_t)b * 2 * HR;
  const uint8_t *row = c.iq + (size_t)b * c.iq_stride;
  for (int k = t; k < 2 * HR; k += blockDim.x) {
    const long long byte = 2 * c.n_rf - 2 * HR + k;  // position in this call's input
    stage[k] = (byte >= 0) ? row[byte] : hrow[2 * HR + byte];
  }
  __syncthreads();
  for (int k = t; k < 2 * HR; k += blockDim.x) hrow[k] = stage[k];
  if (t < 2) c.prev_dst[2 * b + t] = c.prev_src[2 * b + t];
}

}  // namespace sdr
