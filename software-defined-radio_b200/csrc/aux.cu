// aux.cu -- the stages either side of the hot path, device resident:
//
//   sdr_psd / sdr_pipeline_psd   estimatePSD of the reference (src/fourier.cpp:44-126): Bartlett
//                                estimate over 512-sample Hann-windowed segments, as a diagnostics
//                                op on host rows or on an intermediate signal of a pipeline;
//   sdr_deemph_*                 75 us de-emphasis, the output stage the course spec skipped
//                                (doc/3dy4-project-2022.pdf p.6), on the PCM rows in place;
//   sdr_channelizer_*            polyphase analysis bank in FRONT of the receiver: one wideband
//                                uint8 I/Q capture at M x Fs becomes M captures at Fs in HBM, in the
//                                very format sdr_pipeline_process_device consumes, so a batch of
//                                channels is fed from the device instead of across PCIe;
//   sdr_wav_header               RIFF/WAVE header for the PCM the receiver emits.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/sdr_b200.h"
#include "common.cuh"
#include "design.h"

namespace sdr {
int fail(int code, const std::string &msg);

constexpr int PSD_N = 512;        // NFFT, include/dy4.h:27
constexpr int PSD_BINS = PSD_N / 2;
constexpr double PSD_PI = 3.14159265358979323846;   // include/dy4.h:23

// ---------------------------------------------------------------------------------------------
// PSD.  X[s][m] = sum_t (x[s][t] * hann[t]) * exp(-j a(t, m)),  a = float(-2 pi (t m) / 512): the
// reference evaluates the angle in double and stores it in a complex<float> (fourier.cpp:19), so
// the table below is built from the float-rounded angle.  Each (segment, bin) accumulates over t in
// ascending order with separately rounded multiply and add, like fourier.cpp:20.
// CTA = 64 segments x 32 bins, 256 threads, thread = 4 segments x 2 bins.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_psd_dft(const float *__restrict__ x, size_t x_stride, int segs_per_row, int n_seg_total,
                                                  const float2 *__restrict__ tw, const float *__restrict__ hann,
                                                  float inv_norm, float *__restrict__ db /* [n_seg_total][256] */) {
  __shared__ float as[64][33];
  __shared__ float2 ts[32][33];
  const int s0 = blockIdx.x * 64, m0 = blockIdx.y * 32;
  const int t = threadIdx.x, ts_i = t & 15, tm = t >> 4;   // 16 segment groups x 16 bin pairs
  float re[4][2] = {}, im[4][2] = {};
  for (int k0 = 0; k0 < PSD_N; k0 += 32) {
    for (int i = t; i < 64 * 32; i += 256) {
      const int s = s0 + i / 32, k = k0 + i % 32;
      float v = 0.0f;
      if (s < n_seg_total) {
        const int row = s / segs_per_row, seg = s % segs_per_row;
        v = xmul(x[(size_t)row * x_stride + (size_t)seg * PSD_N + k], hann[k]);   // fourier.cpp:79
      }
      as[i / 32][i % 32] = v;
    }
    for (int i = t; i < 32 * 32; i += 256) ts[i / 32][i % 32] = tw[(size_t)(m0 + i / 32) * PSD_N + k0 + i % 32];
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {
      float a[4];
      float2 w[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = as[ts_i * 4 + i][k];
#pragma unroll
      for (int j = 0; j < 2; ++j) w[j] = ts[tm * 2 + j][k];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          re[i][j] = xadd(re[i][j], xmul(a[i], w[j].x));
          im[i][j] = xadd(im[i][j], xmul(a[i], w[j].y));
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int s = s0 + ts_i * 4 + i, m = m0 + tm * 2 + j;
      if (s >= n_seg_total) continue;
      const float mag = sqrtf(xadd(xmul(re[i][j], re[i][j]), xmul(im[i][j], im[i][j])));
      float v = xmul(inv_norm, xmul(mag, mag));   // fourier.cpp:97-98
      v = xmul(2.0f, v);
      db[(size_t)s * PSD_BINS + m] = xmul(10.0f, log10f(v));   // :102-104
    }
}

// fourier.cpp:113-121: the estimate is the MEAN OF THE dB VALUES over the segments, bin by bin.
__global__ void k_psd_mean(const float *__restrict__ db, int segs_per_row, float *__restrict__ psd) {
  const int row = blockIdx.x, m = threadIdx.x;
  float acc = 0.0f;
  for (int s = 0; s < segs_per_row; ++s) acc = xadd(acc, db[((size_t)row * segs_per_row + s) * PSD_BINS + m]);
  psd[(size_t)row * PSD_BINS + m] = segs_per_row ? xdiv(acc, (float)segs_per_row) : 0.0f;
}

struct PsdTables {
  float2 *tw = nullptr;
  float *hann = nullptr;
};
static std::mutex g_psd_mu;
static std::map<int, PsdTables> g_psd_tables;

static int psd_tables(int device, PsdTables *out) {
  std::lock_guard<std::mutex> lk(g_psd_mu);
  auto it = g_psd_tables.find(device);
  if (it == g_psd_tables.end()) {
    std::vector<float2> tw((size_t)PSD_BINS * PSD_N);
    std::vector<float> hann(PSD_N);
    for (int m = 0; m < PSD_BINS; ++m)
      for (int t = 0; t < PSD_N; ++t) {
        const float a = (float)(-2 * PSD_PI * (unsigned)(t * m) / (size_t)PSD_N);   // fourier.cpp:19
        tw[(size_t)m * PSD_N + t] = make_float2((float)std::cos((double)a), (float)std::sin((double)a));
      }
    for (int i = 0; i < PSD_N; ++i) hann[i] = (float)std::pow(std::sin(i * PSD_PI / PSD_N), 2.0);   // :60
    PsdTables t;
    SDR_CUDA(cudaMalloc(&t.tw, tw.size() * sizeof(float2)));
    SDR_CUDA(cudaMalloc(&t.hann, hann.size() * sizeof(float)));
    SDR_CUDA(cudaMemcpy(t.tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    SDR_CUDA(cudaMemcpy(t.hann, hann.data(), hann.size() * sizeof(float), cudaMemcpyHostToDevice));
    it = g_psd_tables.emplace(device, t).first;
  }
  *out = it->second;
  return SDR_OK;
}

// d_x: [rows][x_stride] floats on `device`; d_psd: [rows][256].  Enqueues on `s`; `d_db` scratch of
// rows * segs * 256 floats.
int psd_device(int device, const float *d_x, size_t rows, size_t x_stride, size_t n, float Fs, float *d_db,
               float *d_psd, cudaStream_t s) {
  PsdTables t;
  int rc = psd_tables(device, &t);
  if (rc) return rc;
  const int segs = (int)(n / PSD_N);   // fourier.cpp:70
  if (segs == 0) return fail(SDR_ERR_INVALID, "estimatePSD needs at least 512 samples (fourier.cpp:66-70)");
  const long long total = (long long)rows * segs;
  if (total > 0x7fffffffll / 64) return fail(SDR_ERR_INVALID, "too many PSD segments in one call");
  dim3 grid((unsigned)((total + 63) / 64), PSD_BINS / 32);
  k_psd_dft<<<grid, 256, 0, s>>>(d_x, x_stride, segs, (int)total, t.tw, t.hann, 1 / (Fs * PSD_N / 2), d_db);
  SDR_CUDA(cudaGetLastError());
  k_psd_mean<<<(unsigned)rows, PSD_BINS, 0, s>>>(d_db, segs, d_psd);
  SDR_CUDA(cudaGetLastError());
  return SDR_OK;
}

void psd_freq_axis(float Fs, float *freq) {
  // LinearSpacedArray(freq, Fs/2, 0.0, df), fourier.cpp:36-41,50-54
  const float df = Fs / PSD_N;
  const float N = (Fs / 2 - 0.0f) / df;
  int cnt = 0;
  for (int i = 0; i < N && cnt < PSD_BINS; ++i) freq[cnt++] = 0.0f + i * df;
  for (; cnt < PSD_BINS; ++cnt) freq[cnt] = 0.0f;
}

static int pick_device(int dev) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(SDR_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)");
  }
  if (dev < 0 || dev >= n) return fail(SDR_ERR_INVALID, "device ordinal out of range");
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess || p.major != 10)
    return fail(SDR_ERR_NO_DEVICE, "device is not sm_100 (Blackwell B200); kernels are built for sm_100a only");
  SDR_CUDA(cudaSetDevice(dev));
  return SDR_OK;
}

// ---------------------------------------------------------------------------------------------
// De-emphasis: y[n] = fl(fl(a x[n]) + fl(b y[n-1])), x = pcm / 16384, b = exp(-1/(Fs tau)), a = 1 - b;
// PCM back by truncation like the receiver's own conversion.  One thread per (capture, audio
// channel) walks its row; eight frames per 16-byte load.
// ---------------------------------------------------------------------------------------------
__global__ void k_deemph(int16_t *pcm, size_t stride, size_t n_frames, int channels, int batch, float a, float b,
                         float *state) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= batch * channels) return;
  const int cap = id / channels, ch = id % channels;
  int16_t *row = pcm + (size_t)cap * stride + ch;
  float y = state[id];
  for (size_t i = 0; i < n_frames; ++i) {
    const float x = xmul((float)row[i * channels], 0.00006103515625f);
    y = xadd(xmul(a, x), xmul(b, y));
    row[i * channels] = (int16_t)__float2int_rz(xmul(y, 16384.0f));
  }
  state[id] = y;
}

// ---------------------------------------------------------------------------------------------
// Channeliser.  Output n of channel c:
//   y_c[n] = sum_{r<M} W^{c (M-1-r)} v_r[n],   v_r[n] = sum_{q<T} h[q M + r] x[n M + (M-1) - (q M + r)],
// W = exp(-j 2 pi / M).  A CTA stages the raw bytes of CH_NT outputs (+ history) as float2 and each
// thread makes one output instant for all M channels: M branch sums, then the M-point DFT.
// ---------------------------------------------------------------------------------------------
constexpr int CH_NT = 128;
constexpr int CH_MAXM = 16;

struct ChanArgs {
  const uint8_t *wide;   // [W][wide_stride]
  size_t wide_stride;
  const uint8_t *hist;   // [W][2 * M * T]: the last M*T pairs of the previous call
  const float *h;        // [M * T] prototype
  uint8_t *out;          // [W * M][out_stride]
  size_t out_stride;
  long long n_in;        // wideband pairs in this call
  int n_out;             // outputs per channel in this call (n_in / M)
  int T;
  float gain;
};

template <int M>
__global__ void __launch_bounds__(CH_NT) k_channelize(const ChanArgs a) {
  extern __shared__ float2 xs[];                 // (CH_NT + T) * M pairs
  __shared__ float2 wtab[CH_MAXM];
  const int w = blockIdx.y, n0 = blockIdx.x * CH_NT, t = threadIdx.x;
  const int T = a.T, HW = M * T;
  if (t < M) {
    float s, c;
    sincospif(-2.0f * (float)t / (float)M, &s, &c);
    wtab[t] = make_float2(c, s);
  }
  const uint8_t *row = a.wide + (size_t)w * a.wide_stride;
  const uint8_t *hrow = a.hist + (size_t)w * 2 * HW;
  // pairs [n0*M - HW + M, (n0 + CH_NT)*M) ; tile index j <-> pair n0*M - HW + M + j
  const long long base = (long long)n0 * M - HW + M;
  const int n_pairs = (CH_NT + T - 1) * M;
  for (int j = t; j < n_pairs; j += CH_NT) {
    const long long i = base + j;
    int bi = 128, bq = 128;   // (never read for valid outputs; keeps the arithmetic finite)
    if (i < 0) {
      const long long k = HW + i;
      if (k >= 0) { bi = hrow[2 * k]; bq = hrow[2 * k + 1]; }
    } else if (i < a.n_in) {
      bi = row[2 * i];
      bq = row[2 * i + 1];
    }
    xs[j] = make_float2(((float)bi - 128.0f) * 0.0078125f, ((float)bq - 128.0f) * 0.0078125f);   // iofunc.cpp:133
  }
  __syncthreads();
  const int n = n0 + t;
  if (n >= a.n_out) return;
  // newest pair of output n: n*M + M-1  ->  tile index (t + T - 1) * M + (M - 1) - M + ... = t*M + HW - 1
  const float2 *x_new = xs + (size_t)t * M + HW - 1;
  float2 v[M];
#pragma unroll
  for (int r = 0; r < M; ++r) {
    float sr = 0.0f, si = 0.0f;
    for (int q = 0; q < T; ++q) {
      const float hv = __ldg(a.h + q * M + r);
      const float2 xv = x_new[-(q * M + r)];
      sr = fmaf(hv, xv.x, sr);
      si = fmaf(hv, xv.y, si);
    }
    v[r] = make_float2(sr, si);
  }
#pragma unroll
  for (int c = 0; c < M; ++c) {
    float yr = 0.0f, yi = 0.0f;
#pragma unroll
    for (int r = 0; r < M; ++r) {
      const float2 wv = wtab[(c * (M - 1 - r)) % M];
      yr = fmaf(v[r].x, wv.x, fmaf(-v[r].y, wv.y, yr));
      yi = fmaf(v[r].x, wv.y, fmaf(v[r].y, wv.x, yi));
    }
    // back to the receiver's input format: u8 = 128 + rint(128 g y), clipped (it reads (u8 - 128) / 128)
    const int qi = min(255, max(0, 128 + __float2int_rn(128.0f * a.gain * yr)));
    const int qq = min(255, max(0, 128 + __float2int_rn(128.0f * a.gain * yi)));
    uint8_t *dst = a.out + ((size_t)w * M + c) * a.out_stride + 2 * (size_t)n;
    *reinterpret_cast<uchar2 *>(dst) = make_uchar2((unsigned char)qi, (unsigned char)qq);
  }
}

// ---------------------------------------------------------------------------------------------
// Channeliser, throughput form (the one the C ABI launches; k_channelize above stays as the plain
// statement of the sums, used when the tile would not fit shared memory).
// The bank is M independent T-tap FIRs, one per branch r, each over its OWN phase of the input
// (pairs n M + M-1-r), followed by an M-point DFT across the branches of one output instant.
//  phase 1  thread = (branch r = t % M, run of R = 16 consecutive instants): a sample is loaded
//           once per thread and tap chunk and used for up to 4 taps x 16 instants from registers;
//           (re, im) accumulators are register pairs, one FFMA2 per tap and instant.
//  transpose through shared memory V[r][instant] (padded and skewed: conflict-free both ways)
//  phase 2  thread = instant: reads its M branch sums, u_k = v_{M-1-k}, in-register radix-2 FFT,
//           requantises (clamp, magic-number rounding = rint) and writes one uchar2 per channel row;
//           consecutive lanes are consecutive instants, so every row gets 64 contiguous bytes per warp.
// Staging converts each byte pair to float2 once per CTA (16-byte global loads, 8 pairs each).
// ---------------------------------------------------------------------------------------------
constexpr int CH2_NT = 256, CH2_R = 16;

template <int M>
struct Ch2Geom {
  static constexpr int G = CH2_NT / M;              // instant groups per tile
  static constexpr int NI = G * CH2_R;              // instants per tile
  static constexpr int VPAD = M == 16 ? 0 : M;      // V: a group's R instants are followed by VPAD slots so that the
                                                    // 16/M groups of a half-warp hit different banks
  static constexpr int VROW = G * (CH2_R + VPAD) + 1;   // V row pitch (float2): odd -> lanes r hit distinct banks
  __host__ __device__ static int vcol(int i) { return (i / CH2_R) * (CH2_R + VPAD) + i % CH2_R; }
  __host__ __device__ static int x_pairs(int Tp) { return (NI + Tp - 1) * M; }
  // staged pair p (0 = first pair of the tile's window) -> float2 slot; one group's M pairs of an instant are
  // contiguous, and every R*M pairs the slot index skips M so that the groups of a warp hit different banks
  // ... and inside every chunk of 8 pairs the position is XOR-ed with bits of the chunk number: the
  // staging stores put consecutive LANES on consecutive CHUNKS (64 bytes apart: a 16-way bank conflict
  // without the swizzle, 483 M shared-memory wavefronts against 106 M ideal in profiles/r2/r3q); readers take
  // whole chunks, for which the swizzle is a permutation
  __host__ __device__ static int xpos(int p) {
    const int ps = p ^ ((p >> 4) & 7);
    return ps + (ps / (CH2_R * M)) * M;
  }
  // the transposed branch sums V reuse the staged input's memory (phase 1 has read all of it by then)
  __host__ __device__ static size_t smem_bytes(int Tp) {
    const int last = x_pairs(Tp) + 7;   // the swizzle moves a pair by less than 8 positions
    const size_t xs_b = (size_t)(last + (last / (CH2_R * M)) * M + 8) * sizeof(float2), v_b = (size_t)M * VROW * sizeof(float2);
    return (xs_b > v_b ? xs_b : v_b) + (size_t)((Tp * M + 3) / 4 * 4) * sizeof(float);
  }
};

// in-register DFT: y[c] = sum_k u[k] exp(-2 pi i c k / M), radix-2 decimation in time
template <int M>
__device__ __forceinline__ void ch_dft(float2 (&u)[M]) {
  if constexpr (M == 2) {
    const float2 a = u[0], b = u[1];
    u[0] = make_float2(a.x + b.x, a.y + b.y);
    u[1] = make_float2(a.x - b.x, a.y - b.y);
  } else {
    float2 e[M / 2], o[M / 2];
#pragma unroll
    for (int k = 0; k < M / 2; ++k) { e[k] = u[2 * k]; o[k] = u[2 * k + 1]; }
    ch_dft<M / 2>(e);
    ch_dft<M / 2>(o);
    // exp(-2 pi i k / M), k < M/2, for M up to 16
    constexpr float C16[8] = {1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                              0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f};
    constexpr float S16[8] = {0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f,
                              -1.0f, -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f};
#pragma unroll
    for (int k = 0; k < M / 2; ++k) {
      const float c = C16[k * (16 / M)], sn = S16[k * (16 / M)];
      const float tr = o[k].x * c - o[k].y * sn, ti = o[k].x * sn + o[k].y * c;
      u[k] = make_float2(e[k].x + tr, e[k].y + ti);
      u[k + M / 2] = make_float2(e[k].x - tr, e[k].y - ti);
    }
  }
}
template <>
__device__ __forceinline__ void ch_dft<1>(float2 (&)[1]) {}

template <int M>
__global__ void __launch_bounds__(CH2_NT, 4) k_channelize2(const ChanArgs a, int Tp) {
  using Geo = Ch2Geom<M>;
  constexpr int R = CH2_R, NI = Geo::NI;
  extern __shared__ __align__(16) float2 ch_sm[];
  float *hs = reinterpret_cast<float *>(ch_sm);                    // taps
  float2 *xs = ch_sm + (Tp * M + 3) / 4 * 2;                        // staged input ...
  float2 *V = xs;                                                   // ... and, after phase 1, the branch sums
  const int w = blockIdx.y, n0 = blockIdx.x * NI, t = threadIdx.x;
  const int T = a.T, HW = M * T;
  const uint8_t *row = a.wide + (size_t)w * a.wide_stride;
  const uint8_t *hrow = a.hist + (size_t)w * 2 * HW;
  // taps, scaled by the output gain (the 1/128 of the input conversion and the 128 of the
  // requantisation cancel), zero padded to Tp
  for (int i = t; i < Tp * M; i += CH2_NT) hs[i] = i < HW ? a.h[i] * a.gain : 0.0f;
  // ---- stage pairs [(n0 - (Tp-1)) M, (n0 + NI) M) as (byte - 128) float2 ----
  const long long base = ((long long)n0 - (Tp - 1)) * M;
  const int n_pairs = Geo::x_pairs(Tp);
  const bool aligned = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
  const int off = (int)(((base % 8) + 8) % 8);   // chunks of 8 pairs start on 16-byte boundaries of the row
  for (int j = 8 * t - off; j < n_pairs; j += 8 * CH2_NT) {
    const long long i = base + j;
    float f[16];
    if (aligned && i >= 0 && i + 8 <= a.n_in) {
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(row + 2 * i));
      const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        f[4 * k] = u8_centered(wd[k] & 0xffu);
        f[4 * k + 1] = u8_centered((wd[k] >> 8) & 0xffu);
        f[4 * k + 2] = u8_centered((wd[k] >> 16) & 0xffu);
        f[4 * k + 3] = u8_centered(wd[k] >> 24);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const long long ii = i + k;
        uint32_t bi = 128, bq = 128;   // outside the history / past the end: only zero taps or unused outputs see it
        if (ii < 0) {
          const long long hk = HW + ii;
          if (hk >= 0) { bi = hrow[2 * hk]; bq = hrow[2 * hk + 1]; }
        } else if (ii < a.n_in) {
          bi = row[2 * ii];
          bq = row[2 * ii + 1];
        }
        f[2 * k] = u8_centered(bi);
        f[2 * k + 1] = u8_centered(bq);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (j + k >= 0 && j + k < n_pairs) xs[Geo::xpos(j + k)] = make_float2(f[2 * k], f[2 * k + 1]);
  }
  __syncthreads();
  // ---- phase 1: branch r, instants g*R .. g*R+R-1 of the tile ----
  {
    const int r = t % M, g = t / M;
    f32x2_t acc[R];
#pragma unroll
    for (int j = 0; j < R; ++j) acc[j] = 0ull;
    // sample of instant (tile) i for this branch: staged pair (i + Tp-1) M + M-1-r
    const int col0 = (g * R + Tp - 1) * M + (M - 1 - r);
    // taps in chunks of CH (8 when the padded tap count allows, else 4): the window of R + CH - 1 samples is
    // loaded once per chunk
    auto chunks = [&](auto ch_tag) {
      constexpr int CH = decltype(ch_tag)::value;
      for (int q0 = 0; q0 < Tp; q0 += CH) {
        // window: instants g*R - q0 - (CH-1) .. g*R + R-1 - q0  ->  wv[0 .. R+CH-2]
        f32x2_t wv[R + CH - 1];
#pragma unroll
        for (int k = 0; k < R + CH - 1; ++k)
          wv[k] = *reinterpret_cast<const f32x2_t *>(xs + Geo::xpos(col0 + (k - (CH - 1) - q0) * M));
#pragma unroll
        for (int dq = 0; dq < CH; ++dq) {
          const float hq = hs[(q0 + dq) * M + r];
          const f32x2_t h2 = pack2(hq, hq);
#pragma unroll
          for (int j = 0; j < R; ++j) acc[j] = fma2(acc[j], h2, wv[j + (CH - 1) - dq]);
        }
      }
    };
    if (Tp % 8 == 0) chunks(std::integral_constant<int, 8>{});
    else chunks(std::integral_constant<int, 4>{});
    __syncthreads();   // every thread has read its samples: the staged input's memory becomes V
#pragma unroll
    for (int j = 0; j < R; ++j)
      *reinterpret_cast<f32x2_t *>(V + (size_t)r * Geo::VROW + Geo::vcol(g * R + j)) = acc[j];
  }
  __syncthreads();
  // ---- phase 2: one instant per thread and pass ----
#pragma unroll 1
  for (int e = 0; e < NI / CH2_NT; ++e) {
    const int i = t + e * CH2_NT;
    const int n = n0 + i;
    if (n >= a.n_out) break;
    const int col = Geo::vcol(i);
    float2 u[M];
#pragma unroll
    for (int k = 0; k < M; ++k) u[k] = V[(size_t)(M - 1 - k) * Geo::VROW + col];
    ch_dft<M>(u);
#pragma unroll
    for (int c = 0; c < M; ++c) {
      // u8 = 128 + rint(v), clipped: clamp in float, add 1.5 * 2^23 (the sum's low byte is the
      // two's-complement integer), flip the top bit of the byte for the +128
      const float vr = fminf(fmaxf(u[c].x, -128.0f), 127.0f) + 12582912.0f;
      const float vi = fminf(fmaxf(u[c].y, -128.0f), 127.0f) + 12582912.0f;
      const uint32_t qi = (__float_as_uint(vr) & 0xffu) ^ 0x80u, qq = (__float_as_uint(vi) & 0xffu) ^ 0x80u;
      uint8_t *dst = a.out + ((size_t)w * M + c) * a.out_stride + 2 * (size_t)n;
      *reinterpret_cast<uchar2 *>(dst) = make_uchar2((unsigned char)qi, (unsigned char)qq);
    }
  }
}

__global__ void k_chan_carry(const uint8_t *wide, size_t wide_stride, uint8_t *hist, int HW, long long n_in) {
  extern __shared__ uint8_t stage[];
  const int w = blockIdx.x;
  uint8_t *hrow = hist + (size_t)w * 2 * HW;
  const uint8_t *row = wide + (size_t)w * wide_stride;
  for (int k = threadIdx.x; k < 2 * HW; k += blockDim.x) {
    const long long byte = 2 * n_in - 2 * HW + k;
    stage[k] = byte >= 0 ? row[byte] : hrow[2 * HW + byte];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 2 * HW; k += blockDim.x) hrow[k] = stage[k];
}

}  // namespace sdr

using namespace sdr;

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" int sdr_psd(int device, const float *samples, size_t rows, size_t stride, size_t n, float Fs, float *freq,
                       float *psd) {
  if (!samples || !psd) return fail(SDR_ERR_INVALID, "null argument");
  if (stride < n) return fail(SDR_ERR_INVALID, "stride smaller than the row length");
  int rc = pick_device(device);
  if (rc) return rc;
  const size_t segs = n / PSD_N;
  float *d_x = nullptr, *d_db = nullptr, *d_psd = nullptr;
  auto cleanup = [&] { cudaFree(d_x); cudaFree(d_db); cudaFree(d_psd); };
  if (cudaMalloc(&d_x, rows * n * sizeof(float)) != cudaSuccess || cudaMalloc(&d_db, rows * std::max<size_t>(segs, 1) * PSD_BINS * 4) != cudaSuccess ||
      cudaMalloc(&d_psd, rows * PSD_BINS * 4) != cudaSuccess) {
    cleanup();
    cudaGetLastError();
    return fail(SDR_ERR_NOMEM, "device allocation failed");
  }
  cudaError_t e = cudaMemcpy2D(d_x, n * 4, samples, stride * 4, n * 4, rows, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    rc = psd_device(device, d_x, rows, n, n, Fs, d_db, d_psd, nullptr);
    if (!rc) e = cudaMemcpy(psd, d_psd, rows * PSD_BINS * 4, cudaMemcpyDeviceToHost);
  }
  cleanup();
  if (rc) return rc;
  if (e != cudaSuccess) return cuda_fail(e, "sdr_psd", __FILE__, __LINE__);
  if (freq) psd_freq_axis(Fs, freq);
  return SDR_OK;
}

struct sdr_deemph {
  int device, batch, channels;
  float a, b;
  float *state = nullptr;
};

extern "C" int sdr_deemph_create(int device, int batch, int channels, float Fs, float tau, sdr_deemph **out) {
  if (!out) return fail(SDR_ERR_INVALID, "null argument");
  *out = nullptr;
  if (batch < 1 || channels < 1 || channels > 2 || !(Fs > 0) || !(tau > 0)) return fail(SDR_ERR_INVALID, "bad de-emphasis parameters");
  int rc = pick_device(device);
  if (rc) return rc;
  sdr_deemph *d = new sdr_deemph();
  d->device = device;
  d->batch = batch;
  d->channels = channels;
  d->b = (float)std::exp(-1.0 / ((double)Fs * (double)tau));
  d->a = 1.0f - d->b;
  if (cudaMalloc(&d->state, (size_t)batch * channels * 4) != cudaSuccess || cudaMemset(d->state, 0, (size_t)batch * channels * 4) != cudaSuccess) {
    cudaGetLastError();
    delete d;
    return fail(SDR_ERR_NOMEM, "device allocation failed");
  }
  *out = d;
  return SDR_OK;
}
extern "C" int sdr_deemph_destroy(sdr_deemph *d) {
  if (!d) return SDR_OK;
  cudaSetDevice(d->device);
  cudaFree(d->state);
  delete d;
  return SDR_OK;
}
extern "C" int sdr_deemph_reset(sdr_deemph *d) {
  if (!d) return fail(SDR_ERR_INVALID, "null handle");
  SDR_CUDA(cudaSetDevice(d->device));
  SDR_CUDA(cudaMemset(d->state, 0, (size_t)d->batch * d->channels * 4));
  return SDR_OK;
}
extern "C" int sdr_deemph_process_device(sdr_deemph *d, int16_t *d_pcm, size_t pcm_stride, size_t n_frames, void *stream) {
  if (!d || !d_pcm) return fail(SDR_ERR_INVALID, "null argument");
  if (pcm_stride < n_frames * (size_t)d->channels) return fail(SDR_ERR_INVALID, "pcm_stride too small");
  SDR_CUDA(cudaSetDevice(d->device));
  const int total = d->batch * d->channels;
  k_deemph<<<(total + 63) / 64, 64, 0, (cudaStream_t)stream>>>(d_pcm, pcm_stride, n_frames, d->channels, d->batch, d->a, d->b, d->state);
  SDR_CUDA(cudaGetLastError());
  return SDR_OK;
}
extern "C" int sdr_deemph_process_host(sdr_deemph *d, int16_t *pcm, size_t pcm_stride, size_t n_frames) {
  if (!d || !pcm) return fail(SDR_ERR_INVALID, "null argument");
  SDR_CUDA(cudaSetDevice(d->device));
  const size_t w = n_frames * (size_t)d->channels;
  int16_t *dev = nullptr;
  SDR_CUDA(cudaMalloc(&dev, (size_t)d->batch * w * 2));
  cudaError_t e = cudaMemcpy2D(dev, w * 2, pcm, pcm_stride * 2, w * 2, d->batch, cudaMemcpyHostToDevice);
  int rc = SDR_OK;
  if (e == cudaSuccess) rc = sdr_deemph_process_device(d, dev, w, n_frames, nullptr);
  if (e == cudaSuccess && !rc) e = cudaMemcpy2D(pcm, pcm_stride * 2, dev, w * 2, w * 2, d->batch, cudaMemcpyDeviceToHost);
  cudaFree(dev);
  if (rc) return rc;
  if (e != cudaSuccess) return cuda_fail(e, "sdr_deemph_process_host", __FILE__, __LINE__);
  return SDR_OK;
}

struct sdr_channelizer {
  sdr_channelizer_config cfg;
  std::vector<float> h;
  float *d_h = nullptr;
  uint8_t *d_hist = nullptr;
};

extern "C" int sdr_channelizer_create(const sdr_channelizer_config *cfg, sdr_channelizer **out) {
  if (!cfg || !out) return fail(SDR_ERR_INVALID, "null argument");
  *out = nullptr;
  const int M = cfg->n_channels, T = cfg->taps_per_branch;
  if (!(M == 2 || M == 4 || M == 8 || M == 16)) return fail(SDR_ERR_INVALID, "n_channels must be 2, 4, 8 or 16");
  if (T < 2 || T > 64 || cfg->n_wide < 1 || cfg->n_wide > 65535) return fail(SDR_ERR_INVALID, "bad channeliser parameters");
  int rc = pick_device(cfg->device);
  if (rc) return rc;
  sdr_channelizer *c = new sdr_channelizer();
  c->cfg = *cfg;
  if (!(c->cfg.gain > 0)) c->cfg.gain = 1.0f;
  // prototype: the receiver's own window design (impulseResponseLPF, filter.cpp:103-114) at the
  // wideband rate, cut-off at 80 % of half the channel spacing, M*T taps, unity pass-band gain
  c->h.resize((size_t)M * T);
  design_lpf((float)M, 0.4f, (unsigned short)(M * T), c->h.data());
  const size_t HW = (size_t)M * T;
  if (cudaMalloc(&c->d_h, HW * 4) != cudaSuccess || cudaMalloc(&c->d_hist, (size_t)cfg->n_wide * 2 * HW) != cudaSuccess ||
      cudaMemcpy(c->d_h, c->h.data(), HW * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemset(c->d_hist, 128, (size_t)cfg->n_wide * 2 * HW) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(c->d_h);
    cudaFree(c->d_hist);
    delete c;
    return fail(SDR_ERR_NOMEM, "device allocation failed");
  }
  *out = c;
  return SDR_OK;
}
extern "C" int sdr_channelizer_destroy(sdr_channelizer *c) {
  if (!c) return SDR_OK;
  cudaSetDevice(c->cfg.device);
  cudaFree(c->d_h);
  cudaFree(c->d_hist);
  delete c;
  return SDR_OK;
}
extern "C" int sdr_channelizer_reset(sdr_channelizer *c) {
  if (!c) return fail(SDR_ERR_INVALID, "null handle");
  SDR_CUDA(cudaSetDevice(c->cfg.device));
  SDR_CUDA(cudaMemset(c->d_hist, 128, (size_t)c->cfg.n_wide * 2 * c->h.size()));
  return SDR_OK;
}
extern "C" int sdr_channelizer_prototype(const sdr_channelizer *c, float *h, size_t cap, size_t *n) {
  if (!c || !n) return fail(SDR_ERR_INVALID, "null argument");
  *n = c->h.size();
  if (!h) return SDR_OK;
  if (cap < c->h.size()) return fail(SDR_ERR_CAPACITY, "destination too small");
  std::memcpy(h, c->h.data(), c->h.size() * 4);
  return SDR_OK;
}
extern "C" int sdr_channelizer_process_device(sdr_channelizer *c, const uint8_t *d_wide, size_t wide_stride,
                                              size_t nbytes_wide, uint8_t *d_out, size_t out_stride, void *stream) {
  if (!c || !d_wide || !d_out) return fail(SDR_ERR_INVALID, "null argument");
  const int M = c->cfg.n_channels, T = c->cfg.taps_per_branch;
  if (nbytes_wide % (size_t)(2 * M)) return fail(SDR_ERR_INVALID, "nbytes_wide must be a multiple of 2 * n_channels");
  if (wide_stride < nbytes_wide || out_stride < nbytes_wide / M || (out_stride & 1))
    return fail(SDR_ERR_INVALID, "stride too small");
  if (nbytes_wide == 0) return SDR_OK;
  SDR_CUDA(cudaSetDevice(c->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  ChanArgs a{};
  a.wide = d_wide;
  a.wide_stride = wide_stride;
  a.hist = c->d_hist;
  a.h = c->d_h;
  a.out = d_out;
  a.out_stride = out_stride;
  a.n_in = (long long)(nbytes_wide / 2);
  a.n_out = (int)(a.n_in / M);
  a.T = T;
  a.gain = c->cfg.gain;
  const int Tp = (T + 3) / 4 * 4;
  auto launch2 = [&](auto geom, auto kern) -> bool {   // throughput form, when its tile fits shared memory
    using Geo = decltype(geom);
    const size_t smem2 = Geo::smem_bytes(Tp);
    if (smem2 > 200 * 1024) return false;
    static std::once_flag once[16];
    std::call_once(once[c->cfg.device & 15], [&] {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    });
    dim3 grid2((a.n_out + Geo::NI - 1) / Geo::NI, c->cfg.n_wide);
    kern<<<grid2, CH2_NT, smem2, s>>>(a, Tp);
    return true;
  };
  bool done = false;
  switch (M) {
    case 2: done = launch2(Ch2Geom<2>{}, k_channelize2<2>); break;
    case 4: done = launch2(Ch2Geom<4>{}, k_channelize2<4>); break;
    case 8: done = launch2(Ch2Geom<8>{}, k_channelize2<8>); break;
    default: done = launch2(Ch2Geom<16>{}, k_channelize2<16>); break;
  }
  if (!done) {
    dim3 grid((a.n_out + CH_NT - 1) / CH_NT, c->cfg.n_wide);
    const size_t smem = (size_t)(CH_NT + T) * M * sizeof(float2);
    switch (M) {
      case 2: k_channelize<2><<<grid, CH_NT, smem, s>>>(a); break;
      case 4: k_channelize<4><<<grid, CH_NT, smem, s>>>(a); break;
      case 8: k_channelize<8><<<grid, CH_NT, smem, s>>>(a); break;
      default: k_channelize<16><<<grid, CH_NT, smem, s>>>(a); break;
    }
  }
  SDR_CUDA(cudaGetLastError());
  k_chan_carry<<<c->cfg.n_wide, 128, 2 * (size_t)M * T, s>>>(d_wide, wide_stride, c->d_hist, M * T, a.n_in);
  SDR_CUDA(cudaGetLastError());
  return SDR_OK;
}
extern "C" int sdr_channelizer_process_host(sdr_channelizer *c, const uint8_t *wide, size_t wide_stride, size_t nbytes_wide,
                                            uint8_t *out, size_t out_stride) {
  if (!c || !wide || !out) return fail(SDR_ERR_INVALID, "null argument");
  SDR_CUDA(cudaSetDevice(c->cfg.device));
  const int M = c->cfg.n_channels, W = c->cfg.n_wide;
  const size_t nout = nbytes_wide / M;
  uint8_t *dw = nullptr, *dout = nullptr;
  if (cudaMalloc(&dw, (size_t)W * nbytes_wide) != cudaSuccess || cudaMalloc(&dout, (size_t)W * M * nout) != cudaSuccess) {
    cudaFree(dw);
    cudaGetLastError();
    return fail(SDR_ERR_NOMEM, "device allocation failed");
  }
  cudaError_t e = cudaMemcpy2D(dw, nbytes_wide, wide, wide_stride, nbytes_wide, W, cudaMemcpyHostToDevice);
  int rc = SDR_OK;
  if (e == cudaSuccess) rc = sdr_channelizer_process_device(c, dw, nbytes_wide, nbytes_wide, dout, nout, nullptr);
  if (e == cudaSuccess && !rc) e = cudaMemcpy2D(out, out_stride, dout, nout, nout, (size_t)W * M, cudaMemcpyDeviceToHost);
  cudaFree(dw);
  cudaFree(dout);
  if (rc) return rc;
  if (e != cudaSuccess) return cuda_fail(e, "sdr_channelizer_process_host", __FILE__, __LINE__);
  return SDR_OK;
}

// 44-byte RIFF/WAVE header (PCM, 16 bit) for n_frames frames of `channels` channels.
extern "C" int sdr_wav_header(uint8_t *out44, int sample_rate, int channels, uint64_t n_frames) {
  if (!out44 || sample_rate <= 0 || channels < 1 || channels > 2) return fail(SDR_ERR_INVALID, "bad WAV parameters");
  const uint64_t data = n_frames * (uint64_t)channels * 2;
  const uint32_t data32 = data > 0xffffffffull - 36 ? 0xffffffffu - 36 : (uint32_t)data;   // streaming: clamp
  auto put32 = [&](int at, uint32_t v) { for (int i = 0; i < 4; ++i) out44[at + i] = (uint8_t)(v >> (8 * i)); };
  auto put16 = [&](int at, uint32_t v) { out44[at] = (uint8_t)v; out44[at + 1] = (uint8_t)(v >> 8); };
  std::memcpy(out44, "RIFF", 4);
  put32(4, 36 + data32);
  std::memcpy(out44 + 8, "WAVEfmt ", 8);
  put32(16, 16);
  put16(20, 1);
  put16(22, (uint32_t)channels);
  put32(24, (uint32_t)sample_rate);
  put32(28, (uint32_t)sample_rate * channels * 2);
  put16(32, (uint32_t)channels * 2);
  put16(34, 16);
  std::memcpy(out44 + 36, "data", 4);
  put32(40, data32);
  return SDR_OK;
}
