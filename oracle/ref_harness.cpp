// ref_harness.cpp -- C-ABI harness around the UNMODIFIED reference sources.
//
// TEST INFRASTRUCTURE ONLY.  oracle/Makefile compiles this file together with
// /root/reference/src/filter.cpp and /root/reference/src/iofunc.cpp (where they
// lie, with the reference's own flags: g++ -O3, src/Makefile:4) into
// oracle/_ref/libfmref.so.  No reference source is copied into this repository;
// this file only *calls* the reference's functions (include/filter.h:18-43,
// include/iofunc.h) in the order src/project.cpp does (RF_FrontEnd :80-149,
// RF_MONO :327-381, RF_STEREO :178-309) and exposes the results to the tests.
//
// The binaries `project` / `threadMonoOnly` are not used as the oracle: their
// stdout is text (project) and their tail is racy (exit(1) from the producer).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <streambuf>
#include <vector>

#include "filter.h"  // reference include/filter.h
#include "iofunc.h"  // reference include/iofunc.h
#include "fourier.h" // reference include/fourier.h (estimatePSD: the off-path diagnostics op)

namespace {

// The reference's FastFIR runs one iteration past the end (filter.cpp:166): it
// reads x[x.size()] and writes y[x.size()/decim].  Keep spare capacity behind
// both vectors so that access stays inside owned memory.
constexpr size_t kSlack = 16;

struct membuf : std::streambuf {
  membuf(const uint8_t *p, size_t n) {
    char *c = const_cast<char *>(reinterpret_cast<const char *>(p));
    setg(c, c, c + n);
  }
};

std::vector<float> make(const float *p, size_t n) {
  std::vector<float> v;
  v.reserve(n + kSlack);
  v.assign(p, p + n);
  return v;
}

struct ModeInfo {
  int rf_Fs, if_Fs, audio_Fs, rf_decim, audio_decim, audio_upsamp, block_bytes;
};

bool mode_lookup(int mode, ModeInfo &m) {
  // project.cpp:424-427 (audio_upsamp is 0 there for modes 0/1; 1 here) and :55-57
  switch (mode) {
    case 0: m = {2400000, 240000, 48000, 10, 5, 1, 1024 * 10 * 5 * 2}; return true;
    case 1: m = {1440000, 288000, 48000, 5, 6, 1, 1024 * 5 * 6 * 2}; return true;
    case 2: m = {2400000, 240000, 44100, 10, 800, 147, 7 * 800 * 10 * 2}; return true;
    case 3: m = {960000, 320000, 44100, 3, 3200, 441, 7 * 3200 * 3 * 2}; return true;
  }
  return false;
}

}  // namespace

extern "C" {

struct ref_config {
  int mode, channels, rf_taps, audio_taps, stereo_taps;
};

enum {
  TAP_I_FILT = 0, TAP_Q_FILT, TAP_DEMOD, TAP_ALLPASS, TAP_STEREO_FILT, TAP_CARRIER_FILT,
  TAP_NCO, TAP_MIXER, TAP_AUDIO_FILT, TAP_STEREO_FINAL, TAP_COUNT
};

// ---- host libm probes (what the reference's fmPLL actually calls) --------
void ref_libm_atan2f(const float *y, const float *x, size_t n, float *out) {
  for (size_t i = 0; i < n; i++) out[i] = std::atan2(y[i], x[i]);
}
void ref_libm_sincosf(const float *x, size_t n, float *s, float *c) {
  for (size_t i = 0; i < n; i++) ::sincosf(x[i], &s[i], &c[i]);
}
void ref_libm_cosf(const float *x, size_t n, float *c) {
  for (size_t i = 0; i < n; i++) c[i] = std::cos(x[i]);
}

// ---- primitives -----------------------------------------------------------
void ref_lpf_design(float Fs, float Fc, unsigned short n, float *h) {
  std::vector<float> v;
  impulseResponseLPF(Fs, Fc, n, v);
  std::memcpy(h, v.data(), v.size() * sizeof(float));
}
void ref_bpf_design(float Fs, float Fb, float Fe, unsigned short n, float *h) {
  std::vector<float> v;
  bandPass(Fs, Fb, Fe, n, v);
  std::memcpy(h, v.data(), v.size() * sizeof(float));
}
void ref_u8_to_f32(const uint8_t *raw, size_t n, float *out) {
  membuf mb(raw, n);
  std::streambuf *old = std::cin.rdbuf(&mb);
  std::cin.clear();
  std::vector<float> v(n);
  readStdinBlockData((unsigned)n, 0, v);
  std::cin.rdbuf(old);
  std::cin.clear();
  std::memcpy(out, v.data(), n * sizeof(float));
}
void ref_fir_block(float *y, const float *x, size_t nx, const float *h, size_t nh, float *state) {
  std::vector<float> vy, vx = make(x, nx), vh = make(h, nh), vs = make(state, nh - 1);
  convolveBlockFIR(vy, vx, vh, vs);
  std::memcpy(y, vy.data(), vy.size() * sizeof(float));
  std::memcpy(state, vs.data(), vs.size() * sizeof(float));
}
void ref_fir_decim(float *y, const float *x, size_t nx, const float *h, size_t nh, float *state,
                   unsigned decim) {
  std::vector<float> vy, vx = make(x, nx), vh = make(h, nh), vs = make(state, nh - 1);
  vy.reserve(nx / decim + kSlack);
  convolveBlockFastFIR(vy, vx, vh, vs, decim, false);
  std::memcpy(y, vy.data(), vy.size() * sizeof(float));
  std::memcpy(state, vs.data(), vs.size() * sizeof(float));
}
void ref_fir_resample(float *y, const float *x, size_t nx, const float *h, size_t nh, float *state,
                      unsigned decim, unsigned upsamp) {
  std::vector<float> vy, vx = make(x, nx), vh = make(h, nh), vs = make(state, nh - 1);
  convolveBlockResampleFIR(vy, vx, vh, vs, decim, upsamp, false);
  std::memcpy(y, vy.data(), vy.size() * sizeof(float));
  std::memcpy(state, vs.data(), vs.size() * sizeof(float));
}
void ref_fm_demod(float *out, const float *I, const float *Q, size_t n, float *prev_i,
                  float *prev_q) {
  std::vector<float> vo, vi = make(I, n), vq = make(Q, n);
  fmDemod(vo, vi, vq, *prev_i, *prev_q);
  std::memcpy(out, vo.data(), vo.size() * sizeof(float));
}
void ref_allpass(const float *in, size_t n, float *state, size_t ns, float *out) {
  std::vector<float> vi = make(in, n), vs = make(state, ns), vo;
  allPass(vi, vs, vo);
  std::memcpy(out, vo.data(), vo.size() * sizeof(float));
  std::memcpy(state, vs.data(), ns * sizeof(float));
}
void ref_pll(const float *in, size_t n, float *out, float *state, float freq, float Fs,
             float ncoScale, float phaseAdjust, float normBandwidth) {
  std::vector<float> vi = make(in, n), vo, vs = make(state, 6);
  fmPLL(vi, vo, vs, freq, Fs, ncoScale, phaseAdjust, normBandwidth);
  std::memcpy(out, vo.data(), vo.size() * sizeof(float));
  std::memcpy(state, vs.data(), 6 * sizeof(float));
}
int16_t ref_pcm16(float v) {
  // threadMonoOnly.cpp:188-189
  if (std::isnan(v)) return 0;
  return static_cast<short int>(v * 16384);
}

// ---- whole chain ------------------------------------------------------------
struct ref_chain {
  ref_config cfg;
  ModeInfo mi;
  std::vector<float> rf_h, audio_h, pilot_h, stereo_h;
  std::vector<float> I_state, Q_state, st_mono, st_stereo, st_carrier, st_stereofilt, st_allpass,
      st_pll;
  float prev_i, prev_q;
  std::vector<float> taps[TAP_COUNT];
};

void ref_chain_reset(ref_chain *c) {
  const ref_config &g = c->cfg;
  size_t na = (size_t)g.audio_taps * c->mi.audio_upsamp;
  c->I_state.assign(g.rf_taps - 1, 0.0f);
  c->Q_state.assign(g.rf_taps - 1, 0.0f);
  c->prev_i = c->prev_q = 0.0f;
  // project.cpp:446-458
  c->st_mono.assign(na - 1, 0.0f);
  c->st_stereofilt.assign(na - 1, 0.0f);
  c->st_stereo.assign(g.stereo_taps - 1, 0.0f);
  c->st_carrier.assign(g.stereo_taps - 1, 0.0f);
  c->st_allpass.assign((g.stereo_taps - 1) / 2, 0.0f);
  c->st_pll = {0.0f, 0.0f, 1.0f, 0.0f, 1.0f, 0.0f};
  for (auto &t : c->taps) t.clear();
}

ref_chain *ref_chain_create(const ref_config *cfg) {
  ModeInfo mi;
  if (!mode_lookup(cfg->mode, mi)) return nullptr;
  if (cfg->channels < 1 || cfg->channels > 2) return nullptr;
  size_t na = (size_t)cfg->audio_taps * mi.audio_upsamp;
  if (na > 65535) return nullptr;
  ref_chain *c = new ref_chain();
  c->cfg = *cfg;
  c->mi = mi;
  // project.cpp:50, :165-167, :172-173 (same argument types: ints narrowed to float)
  impulseResponseLPF(mi.rf_Fs, 100000, cfg->rf_taps, c->rf_h);
  impulseResponseLPF(mi.if_Fs * mi.audio_upsamp, 16000, (unsigned short)na, c->audio_h);
  bandPass(mi.if_Fs, 18.5e3, 19.5e3, cfg->stereo_taps, c->pilot_h);
  bandPass(mi.if_Fs, 22e3, 54e3, cfg->stereo_taps, c->stereo_h);
  ref_chain_reset(c);
  return c;
}

void ref_chain_destroy(ref_chain *c) { delete c; }
void ref_chain_clear_taps(ref_chain *c) {
  for (auto &t : c->taps) t.clear();
}
const float *ref_chain_tap(const ref_chain *c, int stage, size_t *n) {
  if (stage < 0 || stage >= TAP_COUNT) { *n = 0; return nullptr; }
  *n = c->taps[stage].size();
  return c->taps[stage].data();
}

static void append(std::vector<float> &dst, const std::vector<float> &src, size_t n) {
  dst.insert(dst.end(), src.begin(), src.begin() + n);
}

size_t ref_chain_process(ref_chain *c, const uint8_t *iq, size_t nbytes, int16_t *pcm, int keep) {
  const ref_config &g = c->cfg;
  const ModeInfo &mi = c->mi;
  const bool resample = g.mode >= 2;
  size_t nblocks = nbytes / (size_t)mi.block_bytes, out = 0;
  membuf mb(iq, nblocks * (size_t)mi.block_bytes);
  std::streambuf *old = std::cin.rdbuf(&mb);
  std::cin.clear();
  auto audio = [&](std::vector<float> &y, const std::vector<float> &x, std::vector<float> &st) {
    y.reserve(x.size() + kSlack);
    if (!resample) convolveBlockFastFIR(y, x, c->audio_h, st, mi.audio_decim, false);
    else convolveBlockResampleFIR(y, x, c->audio_h, st, mi.audio_decim, mi.audio_upsamp, false);
  };
  for (size_t b = 0; b < nblocks; b++) {
    std::vector<float> iq_data(mi.block_bytes);
    readStdinBlockData(mi.block_bytes, (unsigned)b, iq_data);
    std::vector<float> I_in, Q_in;
    I_in.reserve(iq_data.size() / 2 + kSlack);
    Q_in.reserve(iq_data.size() / 2 + kSlack);
    for (size_t k = 0; k < iq_data.size(); k += 2) {
      I_in.push_back(iq_data[k]);
      Q_in.push_back(iq_data[k + 1]);
    }
    std::vector<float> I_filt, Q_filt, demod;
    I_filt.reserve(I_in.size() / mi.rf_decim + kSlack);
    Q_filt.reserve(Q_in.size() / mi.rf_decim + kSlack);
    convolveBlockFastFIR(I_filt, I_in, c->rf_h, c->I_state, mi.rf_decim, false);
    convolveBlockFastFIR(Q_filt, Q_in, c->rf_h, c->Q_state, mi.rf_decim, false);
    fmDemod(demod, I_filt, Q_filt, c->prev_i, c->prev_q);
    demod.reserve(demod.size() + kSlack);
    size_t n_if = demod.size();
    if (keep) {
      append(c->taps[TAP_I_FILT], I_filt, n_if);
      append(c->taps[TAP_Q_FILT], Q_filt, n_if);
      append(c->taps[TAP_DEMOD], demod, n_if);
    }
    if (g.channels == 1) {
      std::vector<float> audio_filt;
      audio(audio_filt, demod, c->st_mono);
      if (keep) append(c->taps[TAP_AUDIO_FILT], audio_filt, audio_filt.size());
      for (float v : audio_filt) pcm[out++] = ref_pcm16(v);
    } else {
      std::vector<float> allp, st_filt, car_filt, audio_filt, nco, mixer, st_final;
      allPass(demod, c->st_allpass, allp);
      allp.reserve(allp.size() + kSlack);
      convolveBlockFIR(st_filt, demod, c->stereo_h, c->st_stereo);
      convolveBlockFIR(car_filt, demod, c->pilot_h, c->st_carrier);
      audio(audio_filt, allp, c->st_mono);
      fmPLL(car_filt, nco, c->st_pll, 19e3, mi.if_Fs, 2.0, 0.0, 0.01);
      mixer.resize(st_filt.size(), 0.0);
      mixer.reserve(mixer.size() + kSlack);
      for (size_t z = 0; z < mixer.size(); z++) mixer[z] = st_filt[z] * nco[z] * 2;
      audio(st_final, mixer, c->st_stereofilt);
      if (keep) {
        append(c->taps[TAP_ALLPASS], allp, n_if);
        append(c->taps[TAP_STEREO_FILT], st_filt, n_if);
        append(c->taps[TAP_CARRIER_FILT], car_filt, n_if);
        append(c->taps[TAP_NCO], nco, n_if);
        append(c->taps[TAP_MIXER], mixer, n_if);
        append(c->taps[TAP_AUDIO_FILT], audio_filt, audio_filt.size());
        append(c->taps[TAP_STEREO_FINAL], st_final, st_final.size());
      }
      for (size_t s = 0; s < st_final.size(); s++) {
        float L = st_final[s] + audio_filt[s];
        float R = audio_filt[s] - st_final[s];
        pcm[out++] = ref_pcm16(L);
        pcm[out++] = ref_pcm16(R);
      }
    }
  }
  std::cin.rdbuf(old);
  std::cin.clear();
  return out;
}

}  // extern "C"

// estimatePSD (src/fourier.cpp:44-126), called as the reference declares it (include/fourier.h:29).
extern "C" int ref_psd(const float *samples, size_t n, float Fs, float *freq, float *psd) {
  std::vector<float> f, p, x(samples, samples + n);
  estimatePSD(f, p, x, Fs);
  for (size_t i = 0; i < 256 && i < f.size(); ++i) freq[i] = f[i];
  for (size_t i = 0; i < 256 && i < p.size(); ++i) psd[i] = p[i];
  return (int)(n / 512);
}
