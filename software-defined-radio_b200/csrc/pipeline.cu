// pipeline.cu -- host side of the C ABI (include/sdr_b200.h): device buffers,
// carried state, kernel selection and launch order for the batched receiver.
//
// Launch order per call (the reference's per-block order, project.cpp:80-149,
// 178-309, 327-381, over an arbitrarily long span instead of one block):
//   mono  : K1 rf+demod -> K3 audio (FIR or resampler) -> carry
//   stereo: K1 rf+demod -> K4 dual band-pass -> K5 PLL -> K6 audio L/R -> carry
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/sdr_b200.h"
#include "design.h"
#include "kernels.cuh"
#include "pipeline_view.h"
#include "rf_tc.cuh"
#include "resample_tc.cuh"
#include <cuda_fp16.h>

namespace sdr {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
  g_err = buf;
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return SDR_ERR_NO_DEVICE;
  if (e == cudaErrorMemoryAllocation) return SDR_ERR_NOMEM;
  return SDR_ERR_CUDA;
}
int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

static bool device_is_sm100(int dev) {
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return false;
  return p.major == 10;
}

static int use_device(int dev) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(SDR_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)");
  }
  if (dev < 0 || dev >= n) return fail(SDR_ERR_INVALID, "device ordinal out of range");
  if (!device_is_sm100(dev))
    return fail(SDR_ERR_NO_DEVICE, "device is not sm_100 (Blackwell B200); kernels are built for sm_100a only");
  SDR_CUDA(cudaSetDevice(dev));
  return SDR_OK;
}

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    release();
    n = count;
    if (!count) return SDR_OK;
    SDR_CUDA(cudaMalloc(&p, count * sizeof(T)));
    return SDR_OK;
  }
  int zero(cudaStream_t s = nullptr) {
    if (p) SDR_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    return SDR_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
};

struct ModeRow {
  int rf_Fs, if_Fs, audio_Fs, rf_decim, audio_decim, audio_upsamp, block_bytes;
};
// project.cpp:424-427 (mode table) and :55-57 (block sizes)
static const ModeRow kModes[4] = {
    {2400000, 240000, 48000, 10, 5, 1, 1024 * 10 * 5 * 2},
    {1440000, 288000, 48000, 5, 6, 1, 1024 * 5 * 6 * 2},
    {2400000, 240000, 44100, 10, 800, 147, 7 * 800 * 10 * 2},
    {960000, 320000, 44100, 3, 3200, 441, 7 * 3200 * 3 * 2},
};

}  // namespace sdr

using namespace sdr;

struct sdr_pipeline {
  sdr_config cfg;
  ModeRow m;
  bool stereo, resample;
  int TA;     // audio taps per phase
  int HR;     // raw I/Q history pairs
  int HD;     // demod history prefix
  int HA;     // stf/nco history prefix
  int delay;  // all-pass delay (stereo), project.cpp:457
  int granule_bytes, if_per_granule, pcm_per_granule;
  int base_granule_bytes, base_if_per_granule, base_pcm_per_granule;  // without a follower stage
  // which audio kernel runs (decided once, at create): modes 0/1 specialised FIR or generic FIRs;
  // modes 2/3 quad resampler (mono), pair resampler (stereo, 101 taps per phase) or the generic one
  enum AudioKernel { AK_FIR, AK_FIR_GENERIC, AK_RS_QUAD, AK_RS_PAIR, AK_RS_GENERIC, AK_RS_TC } audio_kernel = AK_FIR_GENERIC;
  // tensor-core resampler (resample_tc.cuh): period tables, tap tiles, fp16 planes of fm_demod
  RtTables rt_tab{};
  DevBuf<uint8_t> d_rt_tiles;
  DevBuf<uint16_t> xh, xl;
  CUtensorMap map_h{}, map_l{};   // the planes as 2-D tensors [capture][time] for the resampler's TMA copies
  size_t pl_stride = 0;
  int pl_off = 0;
  float rt_out_scale = 0.0f;
  int rs_pitch = 0, rs_rows_cap = 0;   // tile geometry of the chosen resampler
  size_t rs_smem = 0;
  bool scalar_fir = false;             // SDR_VARIANT_SCALAR_FIR: the scalar form of the exact 151-tap FIR kernels
  bool fma_aux = false;                // contract the multiply-adds that do not feed the PLL (FAST, MIXED)
  int n_sm = 148;
  size_t cap_if, cap_audio;  // per-capture capacities of one call
  size_t demod_stride, stf_stride, nco_stride, car_stride, tap_if_stride, tap_audio_stride;
  bool rf_fast, audio_fast, bpf_fast;  // specialised kernels available for these tap counts
  std::vector<float> h_rf, h_audio, h_pilot, h_stereo, h_poly, h_quad;
  DevBuf<float> d_h_rf, d_h_audio, d_h_pilot, d_h_stereo, d_h_poly, d_h_quad;
  int quad_kb = 0;  // rows of one quad table (k_audio_resample_v5)
  DevBuf<int> tc_next_item;  // work counter of the persistent tensor-core front end
  bool tc_counter_armed = false;  // zeroed since its last use (k_carry does it at the end of a call)
  // tensor-core front end (SDR_VARIANT_FAST)
  DevBuf<int8_t> tc_bmat;
  DevBuf<int32_t> tc_hq;
  float tc_scale = 0.0f;
  DevBuf<uint8_t> rf_hist;
  DevBuf<float> prev, prev_new, demod, stf, car, nco, pll_state;
  // optional intermediates (keep_taps or generic-taps path)
  DevBuf<float> iq_filt, t_ifilt, t_qfilt, mix, t_audio, t_stfinal, t_nco, t_allpass;
  bool keep_taps = false;
  size_t last_n_if = 0, last_n_audio = 0;
  uint64_t launches = 0;
  // optional follower stage (RDS, rds.cu): runs at the end of every process call
  sdr::PipelineHook hook = nullptr;
  void *hook_ctx = nullptr;
  // optional per-kernel timing (CUDA events on the launching stream)
  bool profiling = false;
  struct Site {
    std::string name;
    std::vector<cudaEvent_t> ev;  // start,stop pairs
    size_t used = 0;
    double total_ms = 0.0;
    uint64_t count = 0;
  };
  std::vector<Site> sites;
  // host staging for process_host
  uint8_t *pin_in[2] = {nullptr, nullptr};
  int16_t *pin_out[2] = {nullptr, nullptr};
  DevBuf<uint8_t> d_in[2];
  DevBuf<int16_t> d_out[2];
  size_t slice_bytes = 0, slice_pcm = 0;
  cudaStream_t s_copy_in = nullptr, s_compute = nullptr, s_copy_out = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr},
              ev_out[2] = {nullptr, nullptr};
};

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------
// site between prof_begin and check_launch (per host thread: distinct handles may be driven by
// distinct threads)
static thread_local sdr_pipeline::Site *g_open_site = nullptr;
static thread_local cudaStream_t g_open_stream = nullptr;

// Called right before a kernel launch; records the start event when profiling.
static void prof_begin(sdr_pipeline *p, const char *name, cudaStream_t s) {
  g_open_site = nullptr;
  if (!p || !p->profiling) return;
  sdr_pipeline::Site *site = nullptr;
  for (auto &x : p->sites)
    if (x.name == name) site = &x;
  if (!site) {
    p->sites.push_back({});
    site = &p->sites.back();
    site->name = name;
  }
  if (site->used + 2 > site->ev.size()) {
    if (site->ev.size() >= 8192) return;  // bounded: stop sampling this site until drained
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
    site->ev.push_back(a);
    site->ev.push_back(b);
  }
  cudaEventRecord(site->ev[site->used], s);
  g_open_site = site;
  g_open_stream = s;
}

static int check_launch(sdr_pipeline *p, const char *name) {
  cudaError_t e = cudaGetLastError();
  if (g_open_site) {
    cudaEventRecord(g_open_site->ev[g_open_site->used + 1], g_open_stream);
    g_open_site->used += 2;
    g_open_site = nullptr;
  }
  if (e != cudaSuccess) return cuda_fail(e, name, __FILE__, __LINE__);
  if (p) p->launches++;
  return SDR_OK;
}

void sdr_prof_begin(sdr_pipeline *p, const char *name, cudaStream_t s) { prof_begin(p, name, s); }
int sdr_check_launch(sdr_pipeline *p, const char *name) { return check_launch(p, name); }

static int pick_segments(int n_out, int tile_out, int batch, int *outs_per_seg) {
  const int n_tiles = (n_out + tile_out - 1) / tile_out;
  int want = (1776 + batch - 1) / batch;  // ~ 148 SMs x 3 CTAs x 4 waves
  want = std::max(1, std::min(want, n_tiles));
  const int tiles_per_seg = (n_tiles + want - 1) / want;
  *outs_per_seg = tiles_per_seg * tile_out;
  return (n_tiles + tiles_per_seg - 1) / tiles_per_seg;
}

template <int N>
static TapArray<N> make_taps(const std::vector<float> &h) {
  TapArray<N> t;
  std::memset(&t, 0, sizeof t);  // zero padding: see fir_groups
  std::memcpy(t.h, h.data(), sizeof(float) * std::min<size_t>(h.size(), N));
  return t;
}

// Launch with programmatic stream serialisation (see pdl_trigger / pdl_wait in common.cuh): the kernel may
// be scheduled before its predecessor in the stream has finished and waits for it itself.  Only
// kernels that call pdl_wait() before touching their predecessor's output may be launched this way.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ---- K1 dispatch -------------------------------------------------------------
template <int T, int D, int R, int NT, bool MERGE, int ALGO = 0>
static int launch_rf(sdr_pipeline *p, RfArgs a, cudaStream_t s) {
  using Cfg = RfCfg<T, D, R, NT, ALGO>;
  auto kern = k_rf_demod<T, D, R, NT, MERGE, ALGO>;
  static std::once_flag once[16];   // raised once per device, not per launch
  std::call_once(once[p->cfg.device & 15], [&] {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
  });
  int segs = pick_segments(a.n_if, Cfg::TILE_OUT, p->cfg.batch, &a.outs_per_seg);
  dim3 grid(segs, p->cfg.batch);
  prof_begin(p, "k_rf_demod", s);
  kern<<<grid, NT, Cfg::SMEM, s>>>(a, make_taps<Cfg::NTAPS>(p->h_rf));
  return check_launch(p, "k_rf_demod");
}

template <int T, int D, int R, int NT>
static int launch_rf_iq(sdr_pipeline *p, RfArgs a, cudaStream_t s) {
  using Cfg = RfIqCfg<T, D, R, NT>;
  auto kern = k_rf_demod_iq<T, D, R, NT>;
  static std::once_flag once[16];
  std::call_once(once[p->cfg.device & 15], [&] {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
  });
  int segs = pick_segments(a.n_if, Cfg::TILE_OUT, p->cfg.batch, &a.outs_per_seg);
  dim3 grid(segs, p->cfg.batch);
  prof_begin(p, "k_rf_demod", s);
  kern<<<grid, NT, Cfg::SMEM, s>>>(a, make_taps<taps_window(T)>(p->h_rf));
  return check_launch(p, "k_rf_demod");
}

static bool rf_fast_available(int T, int D) {
  return (T == 151 || T == 13) && (D == 10 || D == 5 || D == 3);
}

static int run_rf(sdr_pipeline *p, const RfArgs &a, cudaStream_t s) {
  const int T = p->cfg.rf_taps, D = p->m.rf_decim;
  if (p->cfg.variant == SDR_VARIANT_FAST) {
    RfTcArgs g;
    g.a = a;
    g.bmat = p->tc_bmat.p;
    g.hq = p->tc_hq.p;
    g.scale = p->tc_scale;
    // Persistent CTAs (two per SM) take work items (segment of a capture) from a device counter.
    // Long segments first (a segment start costs an extra barrier and the integer predecessor),
    // a few short ones at the end of every capture so that the last items handed out are small.
    const int n_tiles = (a.n_if + TC_TILE_OUT - 1) / TC_TILE_OUT;
    const int n_cta = 2 * p->n_sm;
    constexpr int items_per_cta = 6, tail_tiles = 8, small_tiles = 4;   // measured on the bench workload
    const int tail = n_tiles >= 4 * tail_tiles ? tail_tiles : 0;          // tiles covered by short segments
    const int n_small = tail ? (tail + small_tiles - 1) / small_tiles : 0;
    const int head = n_tiles - tail;
    const int want = std::max(1, std::min({(n_cta * items_per_cta + p->cfg.batch - 1) / p->cfg.batch, head,
                                           TC_MAX_SEGS - n_small}));
    const int per = (head + want - 1) / want;
    g.segs = 0;
    for (int t = 0; t < head; t += per) g.seg_begin[g.segs++] = t;
    for (int t = head; t < n_tiles; t += small_tiles) g.seg_begin[g.segs++] = t;
    g.seg_begin[g.segs] = n_tiles;
    g.batch = p->cfg.batch;
    dim3 grid(std::min(n_cta, g.segs * g.batch));
    g.next_item = p->tc_next_item.p;
    if (!p->tc_counter_armed &&   // first call, or the previous call failed before its k_carry
        cudaMemsetAsync(g.next_item, 0, sizeof(int), s) != cudaSuccess)
      return fail(SDR_ERR_CUDA, "cudaMemsetAsync");
    p->tc_counter_armed = false;
    prof_begin(p, "k_rf_demod_tc", s);
    if (D == 10) launch_pdl(k_rf_demod_tc<10>, grid, dim3(TC_BLOCK), TcCfg<10>::SMEM, s, g);
    else if (D == 5) launch_pdl(k_rf_demod_tc<5>, grid, dim3(TC_BLOCK), TcCfg<5>::SMEM, s, g);
    else launch_pdl(k_rf_demod_tc<3>, grid, dim3(TC_BLOCK), TcCfg<3>::SMEM, s, g);
    return check_launch(p, "k_rf_demod_tc");
  }
  if (p->rf_fast) {
    // tile shapes from the sweep in DESIGN.md section 4 (two outputs per thread, one merged I/Q loop)
    if (!p->scalar_fir) {   // packed two-rounding multiply-adds (I and Q in one register pair)
      // tile shape: sweep on B200: R = 2 / 4 / 6 / 8 -> 2.14 / 1.99 / 2.33 / 3.20 ms
      if (T == 151 && D == 10) return launch_rf_iq<151, 10, 4, 128>(p, a, s);
      if (T == 151 && D == 5) return launch_rf_iq<151, 5, 4, 128>(p, a, s);
      if (T == 151 && D == 3) return launch_rf_iq<151, 3, 4, 128>(p, a, s);
    }
    if (T == 151 && D == 10) return launch_rf<151, 10, 2, 128, true>(p, a, s);
    if (T == 151 && D == 5) return launch_rf<151, 5, 4, 128, true>(p, a, s);
    if (T == 151 && D == 3) return launch_rf<151, 3, 4, 128, true>(p, a, s);
    if (T == 13 && D == 10) return launch_rf<13, 10, 6, 128, false>(p, a, s);
    if (T == 13 && D == 5) return launch_rf<13, 5, 12, 128, false>(p, a, s);
    if (T == 13 && D == 3) return launch_rf<13, 3, 12, 128, false>(p, a, s);
  }
  RfGenericArgs g;
  g.a = a;
  g.h = p->d_h_rf.p;
  g.T = T;
  g.D = D;
  g.iq_filt = p->iq_filt.p;
  dim3 grid((a.n_if + 127) / 128, p->cfg.batch);
  prof_begin(p, "k_rf_generic", s);
  k_rf_generic<<<grid, 128, T * sizeof(float), s>>>(g);
  int rc = check_launch(p, "k_rf_generic");
  if (rc) return rc;
  prof_begin(p, "k_fm_demod", s);
  k_fm_demod<<<grid, 128, 0, s>>>(p->iq_filt.p, a.n_if, a.demod, a.demod_stride, a.demod_off);
  return check_launch(p, "k_fm_demod");
}

// ---- K3/K6 dispatch (modes 0/1) ----------------------------------------------
template <int T, int D, int R, bool STEREO>
static int launch_audio(sdr_pipeline *p, AudioArgs a, cudaStream_t s) {
  constexpr int NT = 128;
  using Cfg = AudioCfg<T, D, R, NT>;
  constexpr size_t SMEM = (size_t)Cfg::ROW * (STEREO ? 2 : 1) * sizeof(float);
  // FAST / MIXED contract the multiply-adds (the audio filters do not feed the PLL)
  auto kern = p->fma_aux ? k_audio_fir<T, D, R, NT, STEREO, true> : k_audio_fir<T, D, R, NT, STEREO, false>;
  static std::once_flag once[16];   // raised once per device, not per launch
  std::call_once(once[p->cfg.device & 15], [&] {
    cudaFuncSetAttribute(k_audio_fir<T, D, R, NT, STEREO, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    cudaFuncSetAttribute(k_audio_fir<T, D, R, NT, STEREO, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
  });
  int segs = pick_segments(a.n_out, Cfg::TILE_OUT, p->cfg.batch, &a.outs_per_seg);
  dim3 grid(segs, p->cfg.batch);
  prof_begin(p, "k_audio_fir", s);
  if constexpr (STEREO && T == 101) {
    if (!p->fma_aux && !p->scalar_fir) {   // EXACT stereo: the two filters in the lanes of one register pair
      using PCfg = AudioPairCfg<T, D, R, NT>;
      auto pk = k_audio_fir_pair<T, D, R, NT>;
      static std::once_flag once2[16];
      std::call_once(once2[p->cfg.device & 15], [&] {
        cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PCfg::SMEM);
      });
      pk<<<grid, NT, PCfg::SMEM, s>>>(a, make_taps<taps_window(T)>(p->h_audio), 1.0f);
      return check_launch(p, "k_audio_fir");
    }
  }
  kern<<<grid, NT, SMEM, s>>>(a, make_taps<taps_window(T)>(p->h_audio));
  return check_launch(p, "k_audio_fir");
}

static bool audio_fast_available(int T, int D) { return (T == 101 || T == 13) && (D == 5 || D == 6); }

template <bool STEREO>
static int run_audio_fir_fast(sdr_pipeline *p, const AudioArgs &a, cudaStream_t s) {
  const int T = p->TA, D = p->m.audio_decim;
  if (T == 101 && D == 5) return launch_audio<101, 5, 4, STEREO>(p, a, s);
  if (T == 101 && D == 6) return launch_audio<101, 6, 2, STEREO>(p, a, s);
  if (T == 13 && D == 5) return launch_audio<13, 5, 4, STEREO>(p, a, s);
  if (T == 13 && D == 6) return launch_audio<13, 6, 2, STEREO>(p, a, s);
  return fail(SDR_ERR_INVALID, "no specialised audio kernel");
}

// ---- K4 dispatch ---------------------------------------------------------------
template <int T>
static int launch_bpf(sdr_pipeline *p, BpfArgs a, cudaStream_t s) {
  constexpr int NT = 128, R = 12;
  int segs = pick_segments(a.n_if, NT * R, p->cfg.batch, &a.outs_per_seg);
  dim3 grid(segs, p->cfg.batch);
  prof_begin(p, "k_bpf_dual", s);
  if (!p->fma_aux && !p->scalar_fir) {   // EXACT: both filters in the lanes of one register pair
    constexpr int N = taps_groups(T, 1, R);
    TapPairs<N> h2;
    std::memset(&h2, 0, sizeof h2);   // zero padding: see fir_groups
    for (size_t i = 0; i < p->h_stereo.size() && i < (size_t)N; ++i) h2.h[i].x = p->h_stereo[i];
    for (size_t i = 0; i < p->h_pilot.size() && i < (size_t)N; ++i) h2.h[i].y = p->h_pilot[i];
    a.one = 1.0f;
    k_bpf_dual_packed<T, R, NT><<<grid, NT, 0, s>>>(a, h2);
    return check_launch(p, "k_bpf_dual");
  }
  // MIXED: the 22-54 kHz band only reaches the mixer, so its multiply-adds may be contracted; the
  // pilot band feeds the PLL and keeps the reference's two roundings
  auto kern = p->fma_aux ? k_bpf_dual<T, R, NT, true> : k_bpf_dual<T, R, NT, false>;
  kern<<<grid, NT, 0, s>>>(a, make_taps<taps_groups(T, 1, R)>(p->h_stereo),
                           make_taps<taps_groups(T, 1, R)>(p->h_pilot));
  return check_launch(p, "k_bpf_dual");
}

static int run_fir_generic(sdr_pipeline *p, const float *x, size_t xs, int xoff, const float *h,
                           int T, int D, float *y, size_t ys, int yoff, int n_out, int batch,
                           cudaStream_t s) {
  FirGenericArgs g{x, xs, xoff, h, T, D, y, ys, yoff, n_out};
  dim3 grid((n_out + 127) / 128, batch);
  prof_begin(p, "k_fir_generic", s);
  k_fir_generic<<<grid, 128, T * sizeof(float), s>>>(g);
  return check_launch(p, "k_fir_generic");
}

// ---------------------------------------------------------------------------
// C ABI: misc
// ---------------------------------------------------------------------------
extern "C" const char *sdr_version(void) { return "sdr_b200 0.1 (sm_100a)"; }
extern "C" const char *sdr_last_error(void) { return g_err.c_str(); }

extern "C" int sdr_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int d = 0; d < n; ++d) ok += device_is_sm100(d) ? 1 : 0;
  return ok;
}

extern "C" int sdr_mode_lookup(int mode, int channels, sdr_mode_info *out) {
  if (mode < 0 || mode > 3 || channels < 1 || channels > 2 || !out)
    return fail(SDR_ERR_INVALID, "mode must be 0..3 and channels 1|2");
  const ModeRow &m = kModes[mode];
  out->rf_Fs = m.rf_Fs;
  out->if_Fs = m.if_Fs;
  out->audio_Fs = m.audio_Fs;
  out->rf_decim = m.rf_decim;
  out->audio_decim = m.audio_decim;
  out->audio_upsamp = m.audio_upsamp;
  out->block_bytes = m.block_bytes;
  // Smallest span after which every stage is back in phase: the IF count must be a
  // multiple of audio_decim / gcd(audio_decim, audio_upsamp).
  int g = m.audio_decim, r = m.audio_upsamp;
  while (r) { int t = g % r; g = r; r = t; }
  int if_per = m.audio_decim / g;
  out->granule_bytes = if_per * m.rf_decim * 2;
  out->pcm_per_granule = (int)((long long)if_per * m.audio_upsamp / m.audio_decim) * channels;
  return SDR_OK;
}

// ---------------------------------------------------------------------------
// pipeline create / destroy / reset
// ---------------------------------------------------------------------------
static int upload(DevBuf<float> &d, const std::vector<float> &h) {
  int rc = d.alloc(h.size());
  if (rc) return rc;
  SDR_CUDA(cudaMemcpy(d.p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  return SDR_OK;
}

extern "C" int sdr_pipeline_reset(sdr_pipeline *p) {
  if (!p) return fail(SDR_ERR_INVALID, "null pipeline");
  SDR_CUDA(cudaSetDevice(p->cfg.device));
  int rc;
  // raw history = byte 128, i.e. the reference's zero-filled I_state/Q_state (project.cpp:64-65)
  if (p->rf_hist.p) SDR_CUDA(cudaMemset(p->rf_hist.p, 128, p->rf_hist.n));
  if ((rc = p->prev.zero())) return rc;
  if ((rc = p->prev_new.zero())) return rc;
  if ((rc = p->demod.zero())) return rc;
  if ((rc = p->xh.zero()) || (rc = p->xl.zero())) return rc;
  if ((rc = p->stf.zero())) return rc;
  if ((rc = p->car.zero())) return rc;
  if ((rc = p->nco.zero())) return rc;
  if (p->stereo) {
    // project.cpp:458: state_PLL{0,0,1,0,1,0}; nco slot HA is ncoOut[0] of the first block
    std::vector<float> st((size_t)p->cfg.batch * 8, 0.0f);
    for (int b = 0; b < p->cfg.batch; ++b) {
      st[(size_t)b * 8 + 2] = 1.0f;
      st[(size_t)b * 8 + 4] = 1.0f;
    }
    SDR_CUDA(cudaMemcpy(p->pll_state.p, st.data(), st.size() * sizeof(float), cudaMemcpyHostToDevice));
    std::vector<float> ones((size_t)p->cfg.batch, 1.0f);
    SDR_CUDA(cudaMemcpy2D(p->nco.p + p->HA, p->nco_stride * sizeof(float), ones.data(), sizeof(float),
                          sizeof(float), p->cfg.batch, cudaMemcpyHostToDevice));
  }
  p->last_n_if = p->last_n_audio = 0;
  SDR_CUDA(cudaDeviceSynchronize());
  if (p->hook && (rc = p->hook(p->hook_ctx, 1, 0, nullptr))) return rc;
  return SDR_OK;
}

int sdr::pipeline_view(sdr_pipeline *p, DemodView *v) {
  if (!p || !v) return fail(SDR_ERR_INVALID, "null argument");
  v->demod = p->demod.p;
  v->stride = p->demod_stride;
  v->off = p->HD;
  v->batch = p->cfg.batch;
  v->device = p->cfg.device;
  v->mode = p->cfg.mode;
  v->if_Fs = p->m.if_Fs;
  v->rf_decim = p->m.rf_decim;
  v->cap_if = p->cap_if;
  return SDR_OK;
}

int sdr::pipeline_set_hook(sdr_pipeline *p, PipelineHook fn, void *ctx, int granule_bytes) {
  if (!p) return fail(SDR_ERR_INVALID, "null pipeline");
  if (fn && p->hook && p->hook_ctx != ctx) return fail(SDR_ERR_INVALID, "pipeline already has a follower stage");
  // the granule is recomputed from the pipeline's own on every attach and detach
  p->granule_bytes = p->base_granule_bytes;
  p->if_per_granule = p->base_if_per_granule;
  p->pcm_per_granule = p->base_pcm_per_granule;
  if (fn && granule_bytes > 0) {
    long long a = p->granule_bytes, b = granule_bytes;
    while (b) { const long long t = a % b; a = b; b = t; }
    const long long l = (long long)p->granule_bytes / a * granule_bytes;
    const int factor = (int)(l / p->granule_bytes);
    p->granule_bytes = (int)l;
    p->if_per_granule *= factor;
    p->pcm_per_granule *= factor;
  }
  p->hook = fn;
  p->hook_ctx = ctx;
  return SDR_OK;
}

static int alloc_tap_buffers(sdr_pipeline *p) {
  int rc;
  const size_t B = (size_t)p->cfg.batch;
  if (!p->t_ifilt.p) {
    if ((rc = p->t_ifilt.alloc(B * p->tap_if_stride))) return rc;
    if ((rc = p->t_qfilt.alloc(B * p->tap_if_stride))) return rc;
    if ((rc = p->t_audio.alloc(B * p->tap_audio_stride))) return rc;
    if (p->stereo) {
      if ((rc = p->t_stfinal.alloc(B * p->tap_audio_stride))) return rc;
      if ((rc = p->mix.alloc(B * p->stf_stride))) return rc;
      if ((rc = p->t_nco.alloc(B * p->tap_if_stride))) return rc;
      if ((rc = p->t_allpass.alloc(B * p->tap_if_stride))) return rc;
    }
  }
  return SDR_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links cudart only).
static int make_plane_map(CUtensorMap *map, const uint16_t *base, size_t stride_halfs, size_t rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return fail(SDR_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    }
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)stride_halfs, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)stride_halfs * 2};
  const cuuint32_t box[2] = {(cuuint32_t)RT_SLAB, (cuuint32_t)RT_ROWS};   // 32 samples (64 B) x 128 captures
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<uint16_t *>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SDR_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return SDR_OK;
}

// Period tables and tap tiles of the tensor-core resampler (resample_tc.cuh).  Returns false when
// the mode's geometry does not fit the kernel (the quad resampler then stays in charge).
static bool build_resample_tc(sdr_pipeline *p, std::vector<uint8_t> &tiles) {
  const int U = p->m.audio_upsamp, D = p->m.audio_decim, TA = p->TA;
  int g = D, r = U;
  while (r) { const int t = g % r; g = r; r = t; }
  RtTables &t = p->rt_tab;
  t.P_in = D / g;
  t.P_out = U / g;
  if (t.P_in % RT_SLAB) return false;
  t.SP = t.P_in / RT_SLAB;
  t.NBLK = (t.P_out + RT_NB - 1) / RT_NB;
  if (t.SP > RT_MAX_SP || t.NBLK > RT_MAX_BLK) return false;
  auto floordiv32 = [](int v) { return v >= 0 ? v / RT_SLAB : -((-v + RT_SLAB - 1) / RT_SLAB); };
  int n_tiles = 0;
  t.qmin = 0;
  for (int b = 0; b < t.NBLK; ++b) {
    const int j_lo = b * RT_NB, j_hi = std::min(j_lo + RT_NB - 1, t.P_out - 1);
    const int lo = (int)(((long long)j_lo * D) / U) - (TA - 1), hi = (int)(((long long)j_hi * D) / U);
    t.qs[b] = floordiv32(lo);
    t.qe[b] = floordiv32(hi);
    t.tile0[b] = n_tiles;
    n_tiles += t.qe[b] - t.qs[b] + 1;
    t.qmin = std::min(t.qmin, t.qs[b]);
    if (b && t.qs[b] < t.qs[b - 1]) return false;
  }
  if (-t.qmin >= t.SP) return false;   // a block may only reach back into the previous period
  // the schedule: which blocks meet slab position q (at most RT_NACT; a block has long finished when
  // its accumulator slot comes round again: RT_SLOTS = 2 * RT_NACT)
  std::memset(t.sched, 0, sizeof t.sched);
  std::memset(t.any_last, 0, sizeof t.any_last);
  for (int q = 0; q < t.SP; ++q) {
    int n = 0;
    for (int dp = 0; dp < 2; ++dp)
      for (int b = 0; b < t.NBLK; ++b) {
        const int qq = q - dp * t.SP;
        if (qq < t.qs[b] || qq > t.qe[b]) continue;
        if (n == RT_NACT || t.qe[b] - t.qs[b] > 255) return false;
        if (qq == t.qe[b]) t.any_last[q] = 1;
        t.sched[q][n++] = (uint32_t)b | ((uint32_t)(qq - t.qs[b]) << 8) | ((uint32_t)dp << 16) |
                          ((uint32_t)(qq == t.qe[b]) << 17) | (1u << 18);
      }
  }
  // taps with the reference's output gain (filter.cpp:213) and a power-of-two scale folded in
  double hmax = 0.0;
  for (float h : p->h_poly) hmax = std::max(hmax, std::fabs((double)h) * (1.0 + U));
  if (!(hmax > 0.0)) return false;
  int S = 0;
  while (S < 40 && std::ldexp(hmax, S + 1) < 8192.0) ++S;
  while (S > -40 && std::ldexp(hmax, S) >= 8192.0) --S;
  p->rt_out_scale = (float)std::ldexp(1.0, -S);
  // tap tiles, stored per slab position q with the tiles of the blocks active there side by side:
  // [chunk kc][32 rows per entry: hh of its 16 outputs, then hl][8 halfs]
  tiles.assign((size_t)n_tiles * RT_TILE_BYTES, 0);
  std::memset(t.nact, 0, sizeof t.nact);
  std::memset(t.bq_off, 0, sizeof t.bq_off);
  int placed = 0;
  for (int q = 0; q < t.SP; ++q) {
    int n = 0;
    while (n < RT_NACT && (t.sched[q][n] >> 18)) ++n;
    t.nact[q] = (uint32_t)n;
    t.bq_off[q] = (uint32_t)placed;
    uint16_t *blk = reinterpret_cast<uint16_t *>(tiles.data() + (size_t)placed * RT_TILE_BYTES);
    for (int i = 0; i < n; ++i) {
      const uint32_t w = t.sched[q][i];
      const int b = (int)(w & 0xff), js = (int)((w >> 8) & 0xff);
      for (int o = 0; o < RT_NB; ++o) {
        const int j = b * RT_NB + o;
        if (j >= t.P_out) continue;
        const long long m = (long long)j * D;
        const int phase = (int)(m % U), i0 = (int)(m / U);
        for (int k = 0; k < RT_SLAB; ++k) {
          const int in = RT_SLAB * (t.qs[b] + js) + k, tap = i0 - in;
          if (tap < 0 || tap >= TA) continue;
          const double v = std::ldexp((double)p->h_poly[(size_t)phase * TA + tap] * (1.0 + U), S);
          const __half hh = __float2half_rn((float)v);
          const __half hl = __float2half_rn((float)(v - (double)__half2float(hh)));
          const size_t row = (size_t)i * RT_NC + o;
          const size_t at = ((size_t)(k / 8) * (size_t)(RT_NC * n) + row) * 8 + (k % 8);
          blk[at] = __half_as_ushort(hh);
          blk[at + (size_t)RT_NB * 8] = __half_as_ushort(hl);
        }
      }
    }
    placed += n;
  }
  if (placed != n_tiles) return false;
  p->pl_off = -t.qmin * RT_SLAB;
  p->pl_stride = (size_t)round_up((int)(p->pl_off + p->cap_if + RT_SLAB), 8);
  return true;
}

extern "C" int sdr_pipeline_create(const sdr_config *cfg_in, sdr_pipeline **out) {
  if (!cfg_in || !out) return fail(SDR_ERR_INVALID, "null argument");
  *out = nullptr;
  sdr_config cfg_base = *cfg_in;
  const bool scalar_fir = (cfg_base.variant & SDR_VARIANT_SCALAR_FIR) != 0;
  cfg_base.variant &= ~SDR_VARIANT_SCALAR_FIR;
  const sdr_config *cfg = &cfg_base;
  if (cfg->mode < 0 || cfg->mode > 3) return fail(SDR_ERR_INVALID, "mode must be 0..3 (project.cpp:396)");
  if (cfg->channels < 1 || cfg->channels > 2)
    return fail(SDR_ERR_INVALID, "channels must be 1 or 2 (project.cpp:405)");
  if (cfg->batch < 1 || cfg->batch > 65535) return fail(SDR_ERR_INVALID, "batch must be 1..65535");
  if (cfg->rf_taps < 2 || cfg->rf_taps > 1024 || cfg->audio_taps < 2 || cfg->audio_taps > 1024 ||
      cfg->stereo_taps < 3 || cfg->stereo_taps > 1024)
    return fail(SDR_ERR_INVALID, "tap counts out of range");
  if (cfg->variant != SDR_VARIANT_EXACT && cfg->variant != SDR_VARIANT_FAST && cfg->variant != SDR_VARIANT_MIXED)
    return fail(SDR_ERR_INVALID, "unknown variant");
  const ModeRow &m = kModes[cfg->mode];
  if (cfg->variant == SDR_VARIANT_FAST) {
    if (cfg->channels != 1)
      return fail(SDR_ERR_INVALID, "SDR_VARIANT_FAST is mono only: stereo parity needs the bit-exact front end");
    if (cfg->rf_taps > TC_TMAX)
      return fail(SDR_ERR_INVALID, "SDR_VARIANT_FAST supports rf_taps <= 151");
  }
  if ((long long)cfg->audio_taps * m.audio_upsamp > 65535)
    return fail(SDR_ERR_INVALID, "audio_taps*audio_upsamp exceeds unsigned short (filter.h:24)");
  int rc = use_device(cfg->device);
  if (rc) return rc;

  sdr_pipeline *p = new sdr_pipeline();
  p->cfg = *cfg;
  p->m = m;
  p->stereo = cfg->channels == 2;
  p->resample = cfg->mode >= 2;
  p->TA = cfg->audio_taps;
  p->delay = p->stereo ? (cfg->stereo_taps - 1) / 2 : 0;
  p->HR = round_up(cfg->rf_taps - 1 + m.rf_decim, 8);
  if (cfg->variant == SDR_VARIANT_FAST)
    p->HR = std::max(p->HR, m.rf_decim == 10 ? TcCfg<10>::HIST : m.rf_decim == 5 ? TcCfg<5>::HIST : TcCfg<3>::HIST);
  p->HA = round_up(p->TA - 1, 4);
  p->HD = round_up(std::max(p->stereo ? cfg->stereo_taps - 1 : 0, p->TA - 1 + p->delay), 8);  // 32-byte aligned sample 0
  sdr_mode_info mi;
  sdr_mode_lookup(cfg->mode, cfg->channels, &mi);
  p->granule_bytes = p->base_granule_bytes = mi.granule_bytes;
  p->if_per_granule = p->base_if_per_granule = mi.granule_bytes / 2 / m.rf_decim;
  p->pcm_per_granule = p->base_pcm_per_granule = mi.pcm_per_granule;
  p->fma_aux = cfg->variant != SDR_VARIANT_EXACT;
  p->scalar_fir = scalar_fir;
  cudaDeviceGetAttribute(&p->n_sm, cudaDevAttrMultiProcessorCount, cfg->device);
  uint64_t cap_bytes = cfg->max_bytes_per_channel ? cfg->max_bytes_per_channel : (uint64_t)m.block_bytes;
  cap_bytes = (cap_bytes + mi.granule_bytes - 1) / mi.granule_bytes * mi.granule_bytes;
  p->cap_if = cap_bytes / 2 / m.rf_decim;
  p->cap_audio = p->cap_if * m.audio_upsamp / m.audio_decim;
  if (p->cap_if > 0x7fff0000ull / std::max(1, m.audio_upsamp)) {
    delete p;
    return fail(SDR_ERR_INVALID, "max_bytes_per_channel too large for 32-bit sample indices");
  }
  p->rf_fast = rf_fast_available(cfg->rf_taps, m.rf_decim);
  p->audio_fast = !p->resample && audio_fast_available(p->TA, m.audio_decim);
  p->bpf_fast = cfg->stereo_taps == 151 || cfg->stereo_taps == 13;
  // ---- kernels are chosen, and their shared-memory limits raised, once per handle ----
  constexpr size_t kSmemCap = 200 * 1024;
  if (!p->resample) {
    p->audio_kernel = p->audio_fast ? sdr_pipeline::AK_FIR : sdr_pipeline::AK_FIR_GENERIC;
  } else {
    const int U = m.audio_upsamp, D = m.audio_decim;
    // quad form (mono): rows of a tile = inputs spanned by RQ_J outputs + the table (KB + 4 rows) + alignment slack
    const int KB = round_up(p->TA + (U - 1 + 3 * D) / U, 4);
    int pitch = round_up((int)(((long long)(RQ_J - 1) * D) / U) + KB + 4 + 8, 4);
    if ((pitch / 4) % 2 == 0) pitch += 4;  // pitch/4 odd: conflict-free LDS.128 across lanes
    const size_t quad_smem = ((size_t)8 * (KB + 4) * 4 + (size_t)64 * pitch) * sizeof(float) +
                             (size_t)64 * (RQ_J + 2) * sizeof(int16_t);
    // pair / generic forms: rows of the transposed tile = inputs spanned by RS_J outputs + the filter history
    const int rows_cap = (int)(((long long)RS_J * D) / U) + p->TA + 2;
    const int TA4 = (p->TA + 3) & ~3;
    const size_t gen_smem = ((size_t)RS_J * TA4 + (size_t)rows_cap * RS_PITCH * (p->stereo ? 2 : 1)) * sizeof(float) +
                            (size_t)32 * RS_J * (p->stereo ? 2 : 1) * sizeof(int16_t);
    const int du0 = D / U;   // rows between consecutive outputs (floor)
    if (!p->stereo && quad_smem <= kSmemCap) {
      p->audio_kernel = sdr_pipeline::AK_RS_QUAD;
      p->rs_pitch = pitch;
      p->rs_smem = quad_smem;
    } else if (p->stereo && p->TA == 101 && (du0 == 5 || du0 == 7) && gen_smem + 128 <= kSmemCap) {
      p->audio_kernel = sdr_pipeline::AK_RS_PAIR;
      p->rs_rows_cap = rows_cap;
      p->rs_smem = gen_smem + 32 * 2 * sizeof(int16_t);   // padded PCM rows
    } else if (gen_smem <= kSmemCap) {
      p->audio_kernel = sdr_pipeline::AK_RS_GENERIC;
      p->rs_rows_cap = rows_cap;
      p->rs_smem = gen_smem;
    } else {
      delete p;
      return fail(SDR_ERR_INVALID, "audio_taps too large for this mode: no resampler tile fits in shared memory");
    }
  }
  {
    const int big = (int)kSmemCap;
    const cudaFuncAttribute at = cudaFuncAttributeMaxDynamicSharedMemorySize;
    cudaFuncSetAttribute(k_rf_demod_tc<10>, at, (int)TcCfg<10>::SMEM);
    cudaFuncSetAttribute(k_rf_demod_tc<5>, at, (int)TcCfg<5>::SMEM);
    cudaFuncSetAttribute(k_rf_demod_tc<3>, at, (int)TcCfg<3>::SMEM);
    cudaFuncSetAttribute(k_audio_resample_tc<RT_NST, 2>, at, (int)rt_smem(RT_NST));
    cudaFuncSetAttribute(k_audio_resample_v5<true, 2>, at, big);
    cudaFuncSetAttribute(k_audio_resample_v5<false, 2>, at, big);
    cudaFuncSetAttribute(k_audio_resample_v4<true, 101, 5, true>, at, big);
    cudaFuncSetAttribute(k_audio_resample_v4<true, 101, 5, false>, at, big);
    cudaFuncSetAttribute(k_audio_resample_v4<true, 101, 7, true>, at, big);
    cudaFuncSetAttribute(k_audio_resample_v4<true, 101, 7, false>, at, big);
    cudaFuncSetAttribute(k_audio_resample_v2<true, true>, at, big);
    cudaFuncSetAttribute(k_audio_resample_v2<true, false>, at, big);
    cudaFuncSetAttribute(k_audio_resample_v2<false, true>, at, big);
    cudaFuncSetAttribute(k_audio_resample_v2<false, false>, at, big);
    if (cudaGetLastError() != cudaSuccess) {
      delete p;
      return fail(SDR_ERR_CUDA, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed");
    }
  }

  // ---- filter design (host, same arithmetic as the reference; design.cpp) ----
  p->h_rf.resize(cfg->rf_taps);
  design_lpf((float)m.rf_Fs, (float)100000, (unsigned short)cfg->rf_taps, p->h_rf.data());  // project.cpp:50
  const int n_audio_taps = p->TA * m.audio_upsamp;
  p->h_audio.resize(n_audio_taps);
  design_lpf((float)(m.if_Fs * m.audio_upsamp), (float)16000, (unsigned short)n_audio_taps,
             p->h_audio.data());  // project.cpp:165-167
  if (p->stereo) {
    p->h_pilot.resize(cfg->stereo_taps);
    p->h_stereo.resize(cfg->stereo_taps);
    design_bpf((float)m.if_Fs, (float)18.5e3, (float)19.5e3, (unsigned short)cfg->stereo_taps,
               p->h_pilot.data());  // project.cpp:172
    design_bpf((float)m.if_Fs, (float)22e3, (float)54e3, (unsigned short)cfg->stereo_taps,
               p->h_stereo.data());  // project.cpp:173
  }
  // The fused RF kernel scales by 2^-7 after the sum; that is exact only while no
  // product is subnormal.  Designed taps are nowhere near that, but guard anyway.
  for (float h : p->h_rf)
    if (h != 0.0f && std::fabs(h) < 1e-30f) p->rf_fast = false;
  if (p->resample) {
    const int U = m.audio_upsamp;
    p->h_poly.resize((size_t)U * p->TA);
    for (int ph = 0; ph < U; ++ph)
      for (int k = 0; k < p->TA; ++k) p->h_poly[(size_t)ph * p->TA + k] = p->h_audio[ph + (size_t)k * U];
    if (!p->stereo) {
      // quad tables of k_audio_resample_v5: the taps of four consecutive outputs in the order the
      // warp walks the input rows (newest row of the fourth output first)
      const int D = m.audio_decim;
      const int dmax = (U - 1 + 3 * D) / U;
      const int KB = round_up(p->TA + dmax, 4);
      p->quad_kb = KB;
      p->h_quad.assign((size_t)U * KB * 4, 0.0f);
      for (int phi0 = 0; phi0 < U; ++phi0) {
        const int top3 = (phi0 + 3 * D) / U;
        for (int o = 0; o < 4; ++o) {
          const int top = (phi0 + o * D) / U, ph = (phi0 + o * D) % U, d = top3 - top;
          for (int k = 0; k < p->TA; ++k)
            p->h_quad[((size_t)phi0 * KB + (k + d)) * 4 + o] = p->h_poly[(size_t)ph * p->TA + k];
        }
      }
    }
  }

  std::vector<uint8_t> rt_tiles;
  // Tensor-core resampler for both custom-rate modes.  Its cost follows the INPUT rate (one pipeline
  // step per 32-sample slab), the quad resampler's the OUTPUT rate (101 multiply-adds per output).
  // Measured on B200, 1024 captures x 16 blocks: mode 2 0.109 vs 0.245 ms; mode 3 0.350 vs 0.684 ms
  // (the front end's two plane stores instead of one float store cost it 0.03 ms there).
  if (p->resample && !p->stereo && cfg->variant == SDR_VARIANT_FAST && build_resample_tc(p, rt_tiles))
    p->audio_kernel = sdr_pipeline::AK_RS_TC;
  std::vector<int8_t> tc_b;
  std::vector<int32_t> tc_h;
  if (cfg->variant == SDR_VARIANT_FAST) {
    // fixed-point taps: hq = round(h * 2^S), |hq| < 2^30, four balanced base-256 digits
    float hmax = 0.0f;
    for (float h : p->h_rf) hmax = std::max(hmax, std::fabs(h));
    int S = 0;
    while (S < 60 && std::ldexp((double)hmax, S + 1) < 1073741823.0) ++S;
    const int D = m.rf_decim;
    const int Q = (TC_TMAX + D - 1) / D;                 // taps per phase (TcCfg<D>::Q)
    const int K = (16 + Q - 1 + 31) / 32 * 32;           // bytes per A row (TcCfg<D>::K)
    const int BP = TC_N * K;
    const int SHIFT = D == 10 ? 1 : D == 5 ? 2 : 0;       // TcCfg<D>::SHIFT: tap t sits at n = t + SHIFT = D q + ph
    const int BACK = D == 10 ? 32 : TC_FRONT + Q - 1;     // TcCfg<D>::BACK
    tc_h.assign((size_t)D * Q, 0);
    for (int t = 0; t < cfg->rf_taps; ++t)
      tc_h[t + SHIFT] = (int32_t)std::llrint(std::ldexp((double)p->h_rf[t], S));  // round half to even
    p->tc_scale = (float)std::ldexp(1.0, -(S + 7));
    // B tile of phase ph: column 4*delta+d carries digit_d(tap n = D*q+ph) at k = delta + (BACK-FRONT) - q
    tc_b.assign((size_t)D * BP + TC_CORR_BYTES, 0);
    long long digit_sum[TC_ND] = {};
    auto put = [&](int ph, int col, int k, int8_t val) {  // canonical no-swizzle K-major order
      const size_t off = (size_t)ph * BP + ((size_t)(k / 16) * (TC_N / 8) + col / 8) * 128 + (col % 8) * 16 + (k % 16);
      tc_b[off] = val;
    };
    for (int n = 0; n < D * Q; ++n) {
      const int q = n / D, ph = n % D;
      long long v = tc_h[n];
      for (int d = 0; d < TC_ND; ++d) {
        const int digit = (int)(((v + 128) & 255) - 128);
        v = (v - digit) >> 8;
        digit_sum[d] += digit;
        for (int delta = 0; delta < 16; ++delta) put(ph, TC_ND * delta + d, delta + (BACK - TC_FRONT) - q, (int8_t)digit);
      }
    }
    // Offset tile (64 columns x 32): multiplied by an A operand of all 128 it starts every
    // accumulator at -128 * sum_t digit_d(h[t]), i.e. the unsigned samples act as (x - 128).
    // Column (delta, d) holds -digit_sum[d] spread over its 32 entries.
    for (int d = 0; d < TC_ND; ++d) {
      const long long e = -digit_sum[d];
      if (std::llabs(e) > 32 * 127) {
        sdr_pipeline_destroy(p);
        return fail(SDR_ERR_INVALID, "rf taps not representable by the tensor-core front end");
      }
      const int q0 = (int)(e / 32), r = (int)(e - 32 * (long long)q0);
      for (int k = 0; k < 32; ++k) {
        const int val = q0 + (k < std::abs(r) ? (r > 0 ? 1 : -1) : 0);
        for (int delta = 0; delta < 16; ++delta) {
          const int col = TC_ND * delta + d;
          tc_b[(size_t)D * BP + ((size_t)(k / 16) * (TC_N / 8) + col / 8) * 128 + (col % 8) * 16 + (k % 16)] = (int8_t)val;
        }
      }
    }
  }
  const size_t B = (size_t)cfg->batch;
  p->demod_stride = round_up((int)(p->HD + p->cap_if + 8), 8);
  p->stf_stride = round_up((int)(p->HA + p->cap_if + 8), 4);
  p->nco_stride = p->stf_stride;
  p->car_stride = round_up((int)p->cap_if + 4, 4);
  p->tap_if_stride = p->cap_if;
  p->tap_audio_stride = p->cap_audio;
#define TRY(x)            \
  do {                    \
    rc = (x);             \
    if (rc) {             \
      sdr_pipeline_destroy(p); \
      return rc;          \
    }                     \
  } while (0)
  TRY(upload(p->d_h_rf, p->h_rf));
  TRY(upload(p->d_h_audio, p->h_audio));
  if (p->resample) TRY(upload(p->d_h_poly, p->h_poly));
  if (!p->h_quad.empty()) TRY(upload(p->d_h_quad, p->h_quad));
  if (cfg->variant == SDR_VARIANT_FAST) {
    TRY(p->tc_bmat.alloc(tc_b.size()));
    TRY(p->tc_hq.alloc(tc_h.size()));
    TRY(p->tc_next_item.alloc(1));   // zeroed before first use; k_carry re-arms it at the end of every call
    rc = cudaMemcpy(p->tc_bmat.p, tc_b.data(), tc_b.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
                 cudaMemcpy(p->tc_hq.p, tc_h.data(), tc_h.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess
             ? SDR_OK
             : SDR_ERR_CUDA;
    TRY(rc);
  }
  if (p->audio_kernel == sdr_pipeline::AK_RS_TC) {
    TRY(p->d_rt_tiles.alloc(rt_tiles.size()));
    TRY(p->xh.alloc(B * p->pl_stride));
    TRY(p->xl.alloc(B * p->pl_stride));
    TRY(make_plane_map(&p->map_h, p->xh.p, p->pl_stride, B));
    TRY(make_plane_map(&p->map_l, p->xl.p, p->pl_stride, B));
    rc = cudaMemcpy(p->d_rt_tiles.p, rt_tiles.data(), rt_tiles.size(), cudaMemcpyHostToDevice) == cudaSuccess
             ? SDR_OK
             : SDR_ERR_CUDA;
    TRY(rc);
  }
  TRY(p->rf_hist.alloc(B * 2 * p->HR));
  TRY(p->prev.alloc(B * 2));
  TRY(p->prev_new.alloc(B * 2));
  TRY(p->demod.alloc(B * p->demod_stride));
  if (p->stereo) {
    TRY(upload(p->d_h_pilot, p->h_pilot));
    TRY(upload(p->d_h_stereo, p->h_stereo));
    TRY(p->stf.alloc(B * p->stf_stride));
    TRY(p->nco.alloc(B * p->nco_stride));
    TRY(p->car.alloc(B * p->car_stride));
    TRY(p->pll_state.alloc(B * 8));
  }
  if (!p->rf_fast) TRY(p->iq_filt.alloc(B * 2 * (p->cap_if + 1)));
  if (!p->resample && !p->audio_fast) TRY(alloc_tap_buffers(p));
  TRY(sdr_pipeline_reset(p));
#undef TRY
  *out = p;
  return SDR_OK;
}

extern "C" int sdr_pipeline_destroy(sdr_pipeline *p) {
  if (!p) return SDR_OK;
  cudaSetDevice(p->cfg.device);
  cudaDeviceSynchronize();
  for (int i = 0; i < 2; ++i) {
    if (p->pin_in[i]) cudaFreeHost(p->pin_in[i]);
    if (p->pin_out[i]) cudaFreeHost(p->pin_out[i]);
    if (p->ev_in[i]) cudaEventDestroy(p->ev_in[i]);
    if (p->ev_done[i]) cudaEventDestroy(p->ev_done[i]);
    if (p->ev_out[i]) cudaEventDestroy(p->ev_out[i]);
  }
  for (auto &st : p->sites)
    for (auto e : st.ev) cudaEventDestroy(e);
  if (p->s_copy_in) cudaStreamDestroy(p->s_copy_in);
  if (p->s_compute) cudaStreamDestroy(p->s_compute);
  if (p->s_copy_out) cudaStreamDestroy(p->s_copy_out);
  delete p;
  return SDR_OK;
}

extern "C" int sdr_pipeline_keep_taps(sdr_pipeline *p, int keep) {
  if (!p) return fail(SDR_ERR_INVALID, "null pipeline");
  SDR_CUDA(cudaSetDevice(p->cfg.device));
  if (keep) {
    int rc = alloc_tap_buffers(p);
    if (rc) return rc;
  }
  p->keep_taps = keep != 0;
  return SDR_OK;
}

extern "C" int sdr_pipeline_pcm_count(const sdr_pipeline *p, size_t nbytes, size_t *n_pcm) {
  if (!p || !n_pcm) return fail(SDR_ERR_INVALID, "null argument");
  if (nbytes % (size_t)p->granule_bytes) return fail(SDR_ERR_INVALID, "nbytes is not a multiple of the granule");
  *n_pcm = nbytes / (size_t)p->granule_bytes * (size_t)p->pcm_per_granule;
  return SDR_OK;
}

extern "C" int sdr_pipeline_launch_count(sdr_pipeline *p, uint64_t *count, int reset) {
  if (!p) return fail(SDR_ERR_INVALID, "null pipeline");
  if (count) *count = p->launches;
  if (reset) p->launches = 0;
  return SDR_OK;
}

extern "C" int sdr_host_alloc(size_t bytes, void **out) {
  if (!out) return fail(SDR_ERR_INVALID, "null argument");
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(SDR_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)");
  }
  SDR_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
  return SDR_OK;
}

extern "C" int sdr_host_free(void *ptr) {
  if (ptr) SDR_CUDA(cudaFreeHost(ptr));
  return SDR_OK;
}

extern "C" int sdr_pipeline_profile(sdr_pipeline *p, int enable) {
  if (!p) return fail(SDR_ERR_INVALID, "null pipeline");
  p->profiling = enable != 0;
  return SDR_OK;
}

extern "C" int sdr_pipeline_kernel_times(sdr_pipeline *p, int index, char *name, size_t name_cap,
                                         double *total_ms, uint64_t *count, int reset) {
  if (!p) return fail(SDR_ERR_INVALID, "null pipeline");
  if (index < 0 || index >= (int)p->sites.size()) return 1;  // past the end (not an error)
  SDR_CUDA(cudaSetDevice(p->cfg.device));
  sdr_pipeline::Site &st = p->sites[index];
  for (size_t i = 0; i + 1 < st.used; i += 2) {
    SDR_CUDA(cudaEventSynchronize(st.ev[i + 1]));
    float ms = 0.0f;
    SDR_CUDA(cudaEventElapsedTime(&ms, st.ev[i], st.ev[i + 1]));
    st.total_ms += ms;
    st.count++;
  }
  st.used = 0;
  if (name && name_cap) {
    std::strncpy(name, st.name.c_str(), name_cap - 1);
    name[name_cap - 1] = 0;
  }
  if (total_ms) *total_ms = st.total_ms;
  if (count) *count = st.count;
  if (reset) {
    st.total_ms = 0.0;
    st.count = 0;
  }
  return SDR_OK;
}

// ---------------------------------------------------------------------------
// process (device-resident)
// ---------------------------------------------------------------------------
extern "C" int sdr_pipeline_process_device(sdr_pipeline *p, const uint8_t *d_iq, size_t iq_stride,
                                           size_t nbytes, int16_t *d_pcm, size_t pcm_stride,
                                           void *stream) {
  if (!p || !d_iq || !d_pcm) return fail(SDR_ERR_INVALID, "null argument");
  if (nbytes == 0) return SDR_OK;
  if (nbytes % (size_t)p->granule_bytes)
    return fail(SDR_ERR_INVALID, "nbytes_per_channel must be a multiple of the mode's granule");
  if (iq_stride < nbytes) return fail(SDR_ERR_INVALID, "iq_stride smaller than nbytes_per_channel");
  const long long n_rf = (long long)(nbytes / 2);
  const size_t n_if = (size_t)(n_rf / p->m.rf_decim);
  const size_t n_audio = n_if * p->m.audio_upsamp / p->m.audio_decim;
  if (n_if > p->cap_if) return fail(SDR_ERR_CAPACITY, "nbytes_per_channel exceeds max_bytes_per_channel");
  if (pcm_stride < n_audio * (size_t)p->cfg.channels)
    return fail(SDR_ERR_INVALID, "pcm_stride too small");
  SDR_CUDA(cudaSetDevice(p->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const int B = p->cfg.batch;
  const bool taps = p->keep_taps;
  int rc;
  // the follower stage refuses the call before anything is enqueued
  if (p->hook && (rc = p->hook(p->hook_ctx, 2, n_if, s))) return rc;

  // ---- K1 ----
  RfArgs ra{};
  ra.iq = d_iq;
  ra.iq_stride = iq_stride;
  ra.hist = p->rf_hist.p;
  ra.rf_hist_len = p->HR;
  ra.prev_in = p->prev.p;
  ra.prev_out = p->prev_new.p;
  ra.demod = p->demod.p;
  ra.demod_stride = p->demod_stride;
  ra.demod_off = p->HD;
  ra.i_filt = taps ? p->t_ifilt.p : nullptr;
  ra.q_filt = taps ? p->t_qfilt.p : nullptr;
  ra.tap_stride = p->tap_if_stride;
  ra.n_rf = n_rf;
  ra.n_if = (int)n_if;
  ra.one = 1.0f;
  const bool rs_tc = p->audio_kernel == sdr_pipeline::AK_RS_TC;
  ra.write_f32 = 1;
  if (rs_tc) {   // the front end hands fm_demod to the tensor-core resampler as two fp16 planes
    ra.xh = p->xh.p;
    ra.xl = p->xl.p;
    ra.pl_stride = p->pl_stride;
    ra.pl_off = p->pl_off;
    ra.write_f32 = (taps || p->hook) ? 1 : 0;   // the float row only when someone else reads it
  }
  if ((rc = run_rf(p, ra, s))) return rc;

  // ---- stereo: K4 + K5 ----
  if (p->stereo) {
    if (p->bpf_fast) {
      BpfArgs ba{};
      ba.demod = p->demod.p;
      ba.demod_stride = p->demod_stride;
      ba.demod_off = p->HD;
      ba.stf = p->stf.p;
      ba.stf_stride = p->stf_stride;
      ba.hist_off = p->HA;
      ba.car = p->car.p;
      ba.car_stride = p->car_stride;
      ba.n_if = (int)n_if;
      rc = (p->cfg.stereo_taps == 151) ? launch_bpf<151>(p, ba, s) : launch_bpf<13>(p, ba, s);
      if (rc) return rc;
    } else {
      const int T = p->cfg.stereo_taps;
      if ((rc = run_fir_generic(p, p->demod.p, p->demod_stride, p->HD, p->d_h_stereo.p, T, 1,
                                p->stf.p, p->stf_stride, p->HA, (int)n_if, B, s)))
        return rc;
      if ((rc = run_fir_generic(p, p->demod.p, p->demod_stride, p->HD, p->d_h_pilot.p, T, 1,
                                p->car.p, p->car_stride, 0, (int)n_if, B, s)))
        return rc;
    }
    PllArgs pa{};
    pa.in = p->car.p;
    pa.in_stride = p->car_stride;
    pa.out = p->nco.p;
    pa.out_stride = p->nco_stride;
    pa.out_off = p->HA;
    pa.state = p->pll_state.p;
    pa.n = (int)n_if;
    pa.batch = B;
    // project.cpp:237
    pa.freq = (float)19e3;
    pa.Fs = (float)p->m.if_Fs;
    pa.ncoScale = (float)2.0;
    pa.phaseAdjust = (float)0.0;
    pa.normBandwidth = (float)0.01;
    prof_begin(p, "k_pll", s);
    // one lane per capture; up to one warp per scheduler the warps are spread over the SMs (one-warp
    // blocks), beyond that four warps share a block
    const int pll_threads = B > p->n_sm * 4 * 32 ? 32 * PLL_MAX_WARPS : 32;
    k_pll<<<(B + pll_threads - 1) / pll_threads, pll_threads, 0, s>>>(pa);
    if ((rc = check_launch(p, "k_pll"))) return rc;
    prof_begin(p, "k_nco_cos", s);
    k_nco_cos<<<dim3(((unsigned)n_if + 1023) / 1024, B), 256, 0, s>>>(pa);
    if ((rc = check_launch(p, "k_nco_cos"))) return rc;
  }

  // ---- K3 / K6 ----
  AudioArgs aa{};
  aa.demod = p->demod.p;
  aa.demod_stride = p->demod_stride;
  aa.demod_off = p->HD;
  aa.delay = p->delay;
  aa.stf = p->stf.p;
  aa.nco = p->nco.p;
  aa.stf_stride = p->stf_stride;
  aa.nco_stride = p->nco_stride;
  aa.hist_off = p->HA;
  aa.pcm = d_pcm;
  aa.pcm_stride = pcm_stride;
  aa.audio_filt = taps ? p->t_audio.p : nullptr;
  aa.stereo_final = (taps && p->stereo) ? p->t_stfinal.p : nullptr;
  aa.tap_stride = p->tap_audio_stride;
  aa.n_out = (int)n_audio;
  if (p->resample) {
    ResampleArgs g{aa, p->d_h_poly.p, p->m.audio_upsamp, p->m.audio_decim, p->TA};
    const bool fma = p->fma_aux;
    prof_begin(p, "k_audio_resample", s);
    if (rs_tc) {
      ResampleTcArgs t{};
      t.pl_off = p->pl_off;
      t.tiles = p->d_rt_tiles.p;
      t.pcm = d_pcm;
      t.pcm_stride = pcm_stride;
      t.audio_filt = aa.audio_filt;
      t.tap_stride = aa.tap_stride;
      t.out_scale = p->rt_out_scale;
      t.batch = B;
      t.n_periods = (int)(n_if / (size_t)p->rt_tab.P_in);
      const int row_tiles = (B + RT_ROWS - 1) / RT_ROWS;
      const int total_blocks = t.n_periods * p->rt_tab.NBLK;
      t.ctas_per_tile = std::max(1, std::min(total_blocks, (2 * p->n_sm) / row_tiles));
      launch_pdl(k_audio_resample_tc<RT_NST, 2>, dim3(row_tiles * t.ctas_per_tile), dim3(RT_BLOCK), rt_smem(RT_NST), s,
                 t, p->rt_tab, p->map_h, p->map_l);
    } else if (p->audio_kernel == sdr_pipeline::AK_RS_QUAD) {
      ResampleQuadArgs q{aa, p->d_h_quad.p, p->m.audio_upsamp, p->m.audio_decim, p->TA, p->quad_kb, (int)n_if};
      dim3 grid(((int)n_audio + RQ_J - 1) / RQ_J, (B + 63) / 64);
      if (fma) k_audio_resample_v5<true, 2><<<grid, 256, p->rs_smem, s>>>(q, (int)B, p->rs_pitch);
      else k_audio_resample_v5<false, 2><<<grid, 256, p->rs_smem, s>>>(q, (int)B, p->rs_pitch);
    } else {
      dim3 grid(((int)n_audio + RS_J - 1) / RS_J, (B + 31) / 32);
      const int du0 = p->m.audio_decim / p->m.audio_upsamp;
      using ResKernel = void (*)(const ResampleArgs, int, int);
      ResKernel kern;
      if (p->audio_kernel == sdr_pipeline::AK_RS_PAIR)
        kern = du0 == 5 ? (fma ? k_audio_resample_v4<true, 101, 5, true> : k_audio_resample_v4<true, 101, 5, false>)
                        : (fma ? k_audio_resample_v4<true, 101, 7, true> : k_audio_resample_v4<true, 101, 7, false>);
      else if (p->stereo) kern = fma ? k_audio_resample_v2<true, true> : k_audio_resample_v2<true, false>;
      else kern = fma ? k_audio_resample_v2<false, true> : k_audio_resample_v2<false, false>;
      kern<<<grid, RS_NW * 32, p->rs_smem, s>>>(g, B, p->rs_rows_cap);
    }
    if ((rc = check_launch(p, "k_audio_resample"))) return rc;
  } else if (p->audio_fast) {
    rc = p->stereo ? run_audio_fir_fast<true>(p, aa, s) : run_audio_fir_fast<false>(p, aa, s);
    if (rc) return rc;
  } else {
    // generic tap counts: separate FIR launches + PCM pack
    const int T = p->TA, D = p->m.audio_decim;
    if ((rc = run_fir_generic(p, p->demod.p, p->demod_stride, p->HD - p->delay, p->d_h_audio.p, T, D,
                              p->t_audio.p, p->tap_audio_stride, 0, (int)n_audio, B, s)))
      return rc;
    if (p->stereo) {
      dim3 g2(((int)(p->HA + n_if) + 127) / 128, B);
      prof_begin(p, "k_mixer", s);
      k_mixer<<<g2, 128, 0, s>>>(p->stf.p, p->stf_stride, p->nco.p, p->nco_stride, 0,
                                 (int)(p->HA + n_if), p->mix.p, p->stf_stride);
      if ((rc = check_launch(p, "k_mixer"))) return rc;
      if ((rc = run_fir_generic(p, p->mix.p, p->stf_stride, p->HA, p->d_h_audio.p, T, D,
                                p->t_stfinal.p, p->tap_audio_stride, 0, (int)n_audio, B, s)))
        return rc;
    }
    dim3 g3(((int)n_audio + 127) / 128, B);
    prof_begin(p, "k_pcm_pack", s);
    k_pcm_pack<<<g3, 128, 0, s>>>(p->t_audio.p, p->stereo ? p->t_stfinal.p : nullptr,
                                  p->tap_audio_stride, (int)n_audio, d_pcm, pcm_stride);
    if ((rc = check_launch(p, "k_pcm_pack"))) return rc;
  }
  if (taps && p->stereo) {
    // Intermediates that the carry below would clobber are copied out first.
    dim3 g2(((int)(p->HA + n_if) + 127) / 128, B);
    prof_begin(p, "k_mixer", s);
    k_mixer<<<g2, 128, 0, s>>>(p->stf.p, p->stf_stride, p->nco.p, p->nco_stride, 0,
                               (int)(p->HA + n_if), p->mix.p, p->stf_stride);
    if ((rc = check_launch(p, "k_mixer"))) return rc;
    dim3 g3(((int)n_if + 127) / 128, B);
    prof_begin(p, "k_copy_rows", s);
    k_copy_rows<<<g3, 128, 0, s>>>(p->nco.p, p->nco_stride, p->HA, p->t_nco.p, p->tap_if_stride,
                                   (int)n_if);
    if ((rc = check_launch(p, "k_copy_rows"))) return rc;
    prof_begin(p, "k_copy_rows", s);
    k_copy_rows<<<g3, 128, 0, s>>>(p->demod.p, p->demod_stride, p->HD - p->delay, p->t_allpass.p,
                                   p->tap_if_stride, (int)n_if);
    if ((rc = check_launch(p, "k_copy_rows"))) return rc;
  }

  // ---- follower stage (RDS): reads this call's fm_demod before the carry moves on ----
  if (p->hook && (rc = p->hook(p->hook_ctx, 0, n_if, s))) return rc;

  // ---- carry ----
  CarryArgs ca{};
  ca.rows[0] = p->demod.p;
  ca.strides[0] = p->demod_stride;
  ca.src_off[0] = (int)n_if;
  ca.len[0] = p->HD;
  if (p->stereo) {
    ca.rows[1] = p->stf.p;
    ca.strides[1] = p->stf_stride;
    ca.src_off[1] = (int)n_if;
    ca.len[1] = p->HA;
    ca.rows[2] = p->nco.p;
    ca.strides[2] = p->nco_stride;
    ca.src_off[2] = (int)n_if;
    ca.len[2] = p->HA + 1;
  }
  if (rs_tc) {   // planes: rows of halfs moved as 32-bit words (every offset is even)
    ca.rows[3] = reinterpret_cast<float *>(p->xh.p);
    ca.rows[4] = reinterpret_cast<float *>(p->xl.p);
    ca.strides[3] = ca.strides[4] = p->pl_stride / 2;
    ca.src_off[3] = ca.src_off[4] = (int)(n_if / 2);
    ca.len[3] = ca.len[4] = p->pl_off / 2;
  }
  ca.iq = d_iq;
  ca.iq_stride = iq_stride;
  ca.hist = p->rf_hist.p;
  ca.rf_hist_len = p->HR;
  ca.n_rf = n_rf;
  ca.prev_dst = p->prev.p;
  ca.prev_src = p->prev_new.p;
  // When taps are kept the carry would overwrite what sdr_pipeline_tap reads
  // (history-prefixed rows), so the tap accessor accounts for it via last_n_if.
  size_t sh = std::max<size_t>((size_t)std::max({p->HD, p->HA + 1, p->pl_off / 2}) * sizeof(float), (size_t)2 * p->HR);
  ca.work_counter = p->tc_next_item.p;   // nullptr unless the fast variant allocated it
  prof_begin(p, "k_carry", s);
  launch_pdl(k_carry, dim3(B), dim3(128), sh, s, ca);
  if ((rc = check_launch(p, "k_carry"))) return rc;
  p->tc_counter_armed = true;
  p->last_n_if = n_if;
  p->last_n_audio = n_audio;
  return SDR_OK;
}

// ---------------------------------------------------------------------------
// taps
// ---------------------------------------------------------------------------
extern "C" int sdr_pipeline_tap(sdr_pipeline *p, int stage, int channel, float *dst, size_t cap,
                                size_t *n) {
  if (!p || !n) return fail(SDR_ERR_INVALID, "null argument");
  if (channel < 0 || channel >= p->cfg.batch) return fail(SDR_ERR_INVALID, "channel out of range");
  if (!p->keep_taps) return fail(SDR_ERR_INVALID, "call sdr_pipeline_keep_taps(p, 1) before process");
  SDR_CUDA(cudaSetDevice(p->cfg.device));
  const size_t n_if = p->last_n_if, n_audio = p->last_n_audio;
  const float *src = nullptr;
  size_t count = 0;
  const size_t b = (size_t)channel;
  // After the carry, a history-prefixed row holds sample g of the last call at
  // position (prefix + g) for g < n_if - prefix only; the prefix itself now holds
  // the tail.  Rows below are therefore read from the dedicated tap buffers or
  // reassembled from (body, carried tail).
  switch (stage) {
    case SDR_TAP_I_FILT: src = p->t_ifilt.p + b * p->tap_if_stride; count = n_if; break;
    case SDR_TAP_Q_FILT: src = p->t_qfilt.p + b * p->tap_if_stride; count = n_if; break;
    case SDR_TAP_AUDIO_FILT: src = p->t_audio.p + b * p->tap_audio_stride; count = n_audio; break;
    case SDR_TAP_STEREO_FINAL:
      if (!p->stereo) return fail(SDR_ERR_INVALID, "stereo-only tap");
      src = p->t_stfinal.p + b * p->tap_audio_stride; count = n_audio; break;
    case SDR_TAP_CARRIER_FILT:
      if (!p->stereo) return fail(SDR_ERR_INVALID, "stereo-only tap");
      src = p->car.p + b * p->car_stride; count = n_if; break;
    case SDR_TAP_NCO:
      if (!p->stereo) return fail(SDR_ERR_INVALID, "stereo-only tap");
      src = p->t_nco.p + b * p->tap_if_stride; count = n_if; break;
    case SDR_TAP_ALLPASS:
      if (!p->stereo) return fail(SDR_ERR_INVALID, "stereo-only tap");
      src = p->t_allpass.p + b * p->tap_if_stride; count = n_if; break;
    case SDR_TAP_DEMOD: case SDR_TAP_STEREO_FILT: case SDR_TAP_MIXER:
      count = n_if;
      break;
    default:
      return fail(SDR_ERR_INVALID, "unknown tap stage");
  }
  *n = count;
  if (!dst) return SDR_OK;
  if (cap < count) return fail(SDR_ERR_CAPACITY, "tap destination too small");
  SDR_CUDA(cudaDeviceSynchronize());
  if (src) {
    SDR_CUDA(cudaMemcpy(dst, src, count * sizeof(float), cudaMemcpyDeviceToHost));
    return SDR_OK;
  }
  // History-prefixed rows: the carry only rewrites the prefix [0, H), so
  // row[H .. H+n_if) still holds this call's samples.
  const float *row = nullptr;
  size_t H = 0, stride = 0;
  if (stage == SDR_TAP_DEMOD) {
    row = p->demod.p; H = p->HD; stride = p->demod_stride;
  } else if (stage == SDR_TAP_STEREO_FILT) {
    if (!p->stereo) return fail(SDR_ERR_INVALID, "stereo-only tap");
    row = p->stf.p; H = p->HA; stride = p->stf_stride;
  } else {  // mixer: computed by k_mixer into p->mix before the carry
    if (!p->stereo) return fail(SDR_ERR_INVALID, "stereo-only tap");
    row = p->mix.p; H = p->HA; stride = p->stf_stride;
  }
  SDR_CUDA(cudaMemcpy(dst, row + b * stride + H, count * sizeof(float), cudaMemcpyDeviceToHost));
  return SDR_OK;
}

namespace sdr {
int psd_device(int device, const float *d_x, size_t rows, size_t x_stride, size_t n, float Fs, float *d_db,
               float *d_psd, cudaStream_t s);
void psd_freq_axis(float Fs, float *freq);
}  // namespace sdr

// estimatePSD (fourier.cpp:44-126) of intermediate `stage` for every capture of the last call.
extern "C" int sdr_pipeline_psd(sdr_pipeline *p, int stage, float *freq, float *psd) {
  if (!p || !psd) return fail(SDR_ERR_INVALID, "null argument");
  if (!p->keep_taps) return fail(SDR_ERR_INVALID, "call sdr_pipeline_keep_taps(p, 1) before process");
  SDR_CUDA(cudaSetDevice(p->cfg.device));
  const float *rows = nullptr;
  size_t stride = 0, n = 0;
  float Fs = (float)p->m.if_Fs;
  switch (stage) {   // device rows that still hold the whole last call (see sdr_pipeline_tap)
    case SDR_TAP_I_FILT: rows = p->t_ifilt.p; stride = p->tap_if_stride; n = p->last_n_if; break;
    case SDR_TAP_Q_FILT: rows = p->t_qfilt.p; stride = p->tap_if_stride; n = p->last_n_if; break;
    case SDR_TAP_DEMOD: rows = p->demod.p + p->HD; stride = p->demod_stride; n = p->last_n_if; break;
    case SDR_TAP_AUDIO_FILT: rows = p->t_audio.p; stride = p->tap_audio_stride; n = p->last_n_audio; Fs = (float)p->m.audio_Fs; break;
    case SDR_TAP_STEREO_FINAL:
      if (p->stereo) { rows = p->t_stfinal.p; stride = p->tap_audio_stride; n = p->last_n_audio; Fs = (float)p->m.audio_Fs; }
      break;
    case SDR_TAP_CARRIER_FILT: if (p->stereo) { rows = p->car.p; stride = p->car_stride; n = p->last_n_if; } break;
    case SDR_TAP_STEREO_FILT: if (p->stereo) { rows = p->stf.p + p->HA; stride = p->stf_stride; n = p->last_n_if; } break;
    case SDR_TAP_NCO: if (p->stereo) { rows = p->t_nco.p; stride = p->tap_if_stride; n = p->last_n_if; } break;
    case SDR_TAP_MIXER: if (p->stereo) { rows = p->mix.p + p->HA; stride = p->stf_stride; n = p->last_n_if; } break;
    case SDR_TAP_ALLPASS: if (p->stereo) { rows = p->t_allpass.p; stride = p->tap_if_stride; n = p->last_n_if; } break;
    default: return fail(SDR_ERR_INVALID, "unknown tap stage");
  }
  if (!rows) return fail(SDR_ERR_INVALID, "stereo-only tap");
  const size_t B = (size_t)p->cfg.batch, segs = n / 512;
  if (!segs) return fail(SDR_ERR_INVALID, "the last call is shorter than one 512-sample PSD segment");
  DevBuf<float> db, out;
  int rc;
  if ((rc = db.alloc(B * segs * 256)) || (rc = out.alloc(B * 256))) return rc;
  SDR_CUDA(cudaDeviceSynchronize());
  if ((rc = psd_device(p->cfg.device, rows, B, stride, n, Fs, db.p, out.p, nullptr))) return rc;
  SDR_CUDA(cudaMemcpy(psd, out.p, B * 256 * sizeof(float), cudaMemcpyDeviceToHost));
  if (freq) psd_freq_axis(Fs, freq);
  return SDR_OK;
}

extern "C" int sdr_pipeline_copy_state(sdr_pipeline *p, int dst, int src) {
  if (!p) return fail(SDR_ERR_INVALID, "null pipeline");
  if (dst < 0 || src < 0 || dst >= p->cfg.batch || src >= p->cfg.batch)
    return fail(SDR_ERR_INVALID, "channel out of range");
  if (dst == src) return SDR_OK;
  SDR_CUDA(cudaSetDevice(p->cfg.device));
  // the copies below run on the default stream: wait for whatever stream the last process call used
  SDR_CUDA(cudaDeviceSynchronize());
  auto cp = [&](void *base, size_t row_bytes) -> int {
    if (!base) return SDR_OK;
    SDR_CUDA(cudaMemcpy((char *)base + (size_t)dst * row_bytes, (char *)base + (size_t)src * row_bytes,
                        row_bytes, cudaMemcpyDeviceToDevice));
    return SDR_OK;
  };
  int rc;
  if ((rc = cp(p->rf_hist.p, 2 * (size_t)p->HR))) return rc;
  if ((rc = cp(p->prev.p, 2 * sizeof(float)))) return rc;
  // rows are strided: only the prefixes matter, but they sit at row starts
  auto cprow = [&](float *base, size_t stride, size_t len) -> int {
    if (!base) return SDR_OK;
    SDR_CUDA(cudaMemcpy(base + (size_t)dst * stride, base + (size_t)src * stride, len * sizeof(float),
                        cudaMemcpyDeviceToDevice));
    return SDR_OK;
  };
  if ((rc = cprow(p->demod.p, p->demod_stride, p->HD))) return rc;
  if ((rc = cprow(p->stf.p, p->stf_stride, p->HA))) return rc;
  if ((rc = cprow(p->nco.p, p->nco_stride, p->HA + 1))) return rc;
  if ((rc = cprow(p->pll_state.p, 8, 8))) return rc;
  if (p->xh.p) {   // fp16 planes of the tensor-core resampler: history prefix, as 32-bit words
    if ((rc = cprow(reinterpret_cast<float *>(p->xh.p), p->pl_stride / 2, (size_t)p->pl_off / 2))) return rc;
    if ((rc = cprow(reinterpret_cast<float *>(p->xl.p), p->pl_stride / 2, (size_t)p->pl_off / 2))) return rc;
  }
  return SDR_OK;
}

// ---------------------------------------------------------------------------
// process (host buffers): pinned double buffers over three streams replace the
// reference's producer/consumer queue (project.cpp:141-149,181-189,471-496).
// The span is cut along time into slices; slice k+1 is uploaded while slice k
// is computed and slice k-1's PCM is downloaded.
// ---------------------------------------------------------------------------
static int ensure_host_staging(sdr_pipeline *p, size_t slice_bytes) {
  if (p->slice_bytes >= slice_bytes && p->s_compute) return SDR_OK;
  const size_t B = (size_t)p->cfg.batch;
  size_t n_pcm;
  sdr_pipeline_pcm_count(p, slice_bytes, &n_pcm);
  p->slice_bytes = 0;   // nothing usable until every buffer below exists
  for (int i = 0; i < 2; ++i) {
    if (p->pin_in[i]) cudaFreeHost(p->pin_in[i]);
    if (p->pin_out[i]) cudaFreeHost(p->pin_out[i]);
    p->pin_in[i] = nullptr;
    p->pin_out[i] = nullptr;
    SDR_CUDA(cudaMallocHost(&p->pin_in[i], B * slice_bytes));
    SDR_CUDA(cudaMallocHost(&p->pin_out[i], B * n_pcm * sizeof(int16_t)));
    int rc;
    if ((rc = p->d_in[i].alloc(B * slice_bytes))) return rc;
    if ((rc = p->d_out[i].alloc(B * n_pcm))) return rc;
    if (!p->ev_in[i]) {
      SDR_CUDA(cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming));
      SDR_CUDA(cudaEventCreateWithFlags(&p->ev_done[i], cudaEventDisableTiming));
      SDR_CUDA(cudaEventCreateWithFlags(&p->ev_out[i], cudaEventDisableTiming));
    }
  }
  if (!p->s_compute) {
    SDR_CUDA(cudaStreamCreateWithFlags(&p->s_copy_in, cudaStreamNonBlocking));
    SDR_CUDA(cudaStreamCreateWithFlags(&p->s_compute, cudaStreamNonBlocking));
    SDR_CUDA(cudaStreamCreateWithFlags(&p->s_copy_out, cudaStreamNonBlocking));
  }
  p->slice_bytes = slice_bytes;
  p->slice_pcm = n_pcm;
  return SDR_OK;
}

static bool is_pinned(const void *ptr) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

extern "C" int sdr_pipeline_process_host(sdr_pipeline *p, const uint8_t *iq, size_t iq_stride,
                                         size_t nbytes, int16_t *pcm, size_t pcm_stride) {
  if (!p || !iq || !pcm) return fail(SDR_ERR_INVALID, "null argument");
  if (nbytes % (size_t)p->granule_bytes)
    return fail(SDR_ERR_INVALID, "nbytes_per_channel must be a multiple of the mode's granule");
  if (nbytes == 0) return SDR_OK;
  SDR_CUDA(cudaSetDevice(p->cfg.device));
  const size_t B = (size_t)p->cfg.batch;
  const size_t gran = (size_t)p->granule_bytes;
  // the follower stage refuses the whole call before its first slice is enqueued
  if (p->hook) {
    const int hrc = p->hook(p->hook_ctx, 4, nbytes / 2 / (size_t)p->m.rf_decim, nullptr);
    if (hrc) return hrc;
  }
  // slice: <= capacity, and about 64 MiB over the whole batch
  size_t cap_bytes = p->cap_if * p->m.rf_decim * 2;
  size_t slice = std::max<size_t>(gran, ((64ull << 20) / B) / gran * gran);
  slice = std::min(slice, cap_bytes / gran * gran);
  slice = std::min(slice, nbytes);
  int rc = ensure_host_staging(p, slice);
  if (rc) return rc;
  slice = std::min(p->slice_bytes, nbytes) / gran * gran;
  const bool in_pinned = is_pinned(iq), out_pinned = is_pinned(pcm);
  const size_t n_slices = (nbytes + slice - 1) / slice;
  size_t pcm_done = 0;
  size_t pcm_off[2] = {0, 0}, pcm_cnt[2] = {0, 0};
  auto drain = [&](int buf) -> int {  // wait for buffer's D2H and hand PCM to the caller
    SDR_CUDA(cudaEventSynchronize(p->ev_out[buf]));
    if (!out_pinned && pcm_cnt[buf]) {
      for (size_t b = 0; b < B; ++b)
        std::memcpy(pcm + b * pcm_stride + pcm_off[buf], p->pin_out[buf] + b * pcm_cnt[buf],
                    pcm_cnt[buf] * sizeof(int16_t));
    }
    pcm_cnt[buf] = 0;
    return SDR_OK;
  };
  for (size_t k = 0; k < n_slices; ++k) {
    const int buf = (int)(k & 1);
    const size_t off = k * slice;
    const size_t len = std::min(slice, nbytes - off);
    size_t n_pcm;
    sdr_pipeline_pcm_count(p, len, &n_pcm);
    if (k >= 2) {
      // this buffer pair was last used by slice k-2: its PCM must have reached the
      // caller, its compute must be finished before the device input is overwritten
      if ((rc = drain(buf))) return rc;
      SDR_CUDA(cudaStreamWaitEvent(p->s_copy_in, p->ev_done[buf], 0));
      if (!in_pinned) SDR_CUDA(cudaEventSynchronize(p->ev_in[buf]));
    }
    // ---- upload ----
    if (in_pinned) {
      SDR_CUDA(cudaMemcpy2DAsync(p->d_in[buf].p, len, iq + off, iq_stride, len, B,
                                 cudaMemcpyHostToDevice, p->s_copy_in));
    } else {
      for (size_t b = 0; b < B; ++b) std::memcpy(p->pin_in[buf] + b * len, iq + b * iq_stride + off, len);
      SDR_CUDA(cudaMemcpyAsync(p->d_in[buf].p, p->pin_in[buf], B * len, cudaMemcpyHostToDevice,
                               p->s_copy_in));
    }
    SDR_CUDA(cudaEventRecord(p->ev_in[buf], p->s_copy_in));
    // ---- compute (after its input arrived and slice k-2's PCM left d_out[buf]) ----
    SDR_CUDA(cudaStreamWaitEvent(p->s_compute, p->ev_in[buf], 0));
    if (k >= 2) SDR_CUDA(cudaStreamWaitEvent(p->s_compute, p->ev_out[buf], 0));
    if ((rc = sdr_pipeline_process_device(p, p->d_in[buf].p, len, len, p->d_out[buf].p, n_pcm,
                                          p->s_compute)))
      return rc;
    SDR_CUDA(cudaEventRecord(p->ev_done[buf], p->s_compute));
    // ---- download ----
    SDR_CUDA(cudaStreamWaitEvent(p->s_copy_out, p->ev_done[buf], 0));
    if (out_pinned) {
      SDR_CUDA(cudaMemcpy2DAsync(pcm + pcm_done, pcm_stride * sizeof(int16_t), p->d_out[buf].p,
                                 n_pcm * sizeof(int16_t), n_pcm * sizeof(int16_t), B,
                                 cudaMemcpyDeviceToHost, p->s_copy_out));
    } else {
      SDR_CUDA(cudaMemcpyAsync(p->pin_out[buf], p->d_out[buf].p, B * n_pcm * sizeof(int16_t),
                               cudaMemcpyDeviceToHost, p->s_copy_out));
    }
    SDR_CUDA(cudaEventRecord(p->ev_out[buf], p->s_copy_out));
    pcm_off[buf] = pcm_done;
    pcm_cnt[buf] = n_pcm;
    pcm_done += n_pcm;
  }
  for (size_t k = (n_slices >= 2 ? n_slices - 2 : 0); k < n_slices; ++k)
    if ((rc = drain((int)(k & 1)))) return rc;
  SDR_CUDA(cudaStreamSynchronize(p->s_compute));
  return SDR_OK;
}

#ifdef SDR_TC_TRACE
extern "C" int sdr_debug_tc_times(long long *begin, long long *end, int n) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(begin, sdr::g_tc_begin, (size_t)n * sizeof(long long)) == cudaSuccess &&
                 cudaMemcpyFromSymbol(end, sdr::g_tc_end, (size_t)n * sizeof(long long)) == cudaSuccess
             ? 0
             : 1;
}
extern "C" int sdr_debug_tc_trace(long long *out, int n) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, sdr::g_tc_trace, (size_t)n * sizeof(long long)) == cudaSuccess ? 0 : 1;
}
#endif
