// shim_parity.cpp -- exercises include/dropin/filter.h (the reference's C++ API on the GPU)
// exactly the way src/project.cpp uses it, and checks every result bit for bit against the
// CPU oracle (oracle/fm_oracle.h).  Built and run by tests/test_gpu_cpp.py.
#include <cstdio>
#include <cstring>
#include <random>
#include <stdexcept>
#include <vector>

#include "dropin/filter.h"
#include "fm_oracle.h"

static int failures = 0;

static void expect_same(const char *what, const std::vector<float> &a, const std::vector<float> &b) {
  bool ok = a.size() == b.size() && (a.empty() || std::memcmp(a.data(), b.data(), a.size() * 4) == 0);
  std::printf("%-34s %s (n=%zu)\n", what, ok ? "ok" : "MISMATCH", a.size());
  if (!ok) ++failures;
}

int main() {
  std::mt19937 rng(3);
  std::normal_distribution<float> nd(0.0f, 1.0f);
  auto randv = [&](size_t n) { std::vector<float> v(n); for (auto &x : v) x = nd(rng); return v; };

  // design: same call sites as project.cpp:50,165,172-173
  std::vector<float> rf, au, pilot, want;
  impulseResponseLPF(2400000, 100000, 151, rf);
  want.assign(151, 0); orc_lpf_design(2400000, 100000, 151, want.data());
  expect_same("impulseResponseLPF rf", rf, want);
  bandPass(240000, 18.5e3, 19.5e3, 151, pilot);
  want.assign(151, 0); orc_bpf_design(240000, 18.5e3f, 19.5e3f, 151, want.data());
  expect_same("bandPass pilot", pilot, want);
  impulseResponseLPF(240000, 16000, 101, au);

  // front end of one block: two FastFIRs + fmDemod with carried state, three blocks
  std::vector<float> I_state(150, 0.0f), Q_state(150, 0.0f), oI(150, 0.0f), oQ(150, 0.0f);
  float prev_i = 0, prev_q = 0, opi = 0, opq = 0;
  std::vector<float> st_mono(100, 0.0f), ost_mono(100, 0.0f), st_car(150, 0.0f), ost_car(150, 0.0f);
  std::vector<float> st_ap(75, 0.0f), ost_ap(75, 0.0f), pll{0, 0, 1, 0, 1, 0}, opll = pll;
  for (int blk = 0; blk < 3; ++blk) {
    std::vector<float> I = randv(51200), Q = randv(51200), If, Qf, dem, aud, car, ap, nco;
    convolveBlockFastFIR(If, I, rf, I_state, 10, false);
    convolveBlockFastFIR(Qf, Q, rf, Q_state, 10, false);
    fmDemod(dem, If, Qf, prev_i, prev_q);
    std::vector<float> wIf(5120), wQf(5120), wdem(5120);
    orc_fir_decim(wIf.data(), I.data(), I.size(), rf.data(), rf.size(), oI.data(), 10);
    orc_fir_decim(wQf.data(), Q.data(), Q.size(), rf.data(), rf.size(), oQ.data(), 10);
    orc_fm_demod(wdem.data(), wIf.data(), wQf.data(), 5120, &opi, &opq);
    expect_same("convolveBlockFastFIR I", If, wIf);
    expect_same("fmDemod", dem, wdem);
    expect_same("I_state", I_state, oI);
    convolveBlockFastFIR(aud, dem, au, st_mono, 5, false);
    std::vector<float> waud(1024);
    orc_fir_decim(waud.data(), wdem.data(), 5120, au.data(), au.size(), ost_mono.data(), 5);
    expect_same("convolveBlockFastFIR audio", aud, waud);
    convolveBlockFIR(car, dem, pilot, st_car);
    std::vector<float> wcar(5120);
    orc_fir_block(wcar.data(), wdem.data(), 5120, pilot.data(), pilot.size(), ost_car.data());
    expect_same("convolveBlockFIR pilot", car, wcar);
    allPass(dem, st_ap, ap);
    std::vector<float> wap(5120);
    orc_allpass(wdem.data(), 5120, ost_ap.data(), 75, wap.data());
    expect_same("allPass", ap, wap);
    fmPLL(car, nco, pll, 19e3, 240000, 2.0, 0.0, 0.01);
    std::vector<float> wnco(5121);
    orc_pll(wcar.data(), 5120, wnco.data(), opll.data(), 19e3f, 240000, 2.0f, 0.0f, 0.01f);
    expect_same("fmPLL ncoOut", nco, wnco);
    expect_same("fmPLL state", pll, opll);
  }
  // mode-2 resampler with the reference's zero-stuffed state
  {
    std::vector<float> h;
    impulseResponseLPF(240000 * 147, 16000, 101 * 147, h);
    std::vector<float> st(h.size() - 1, 0.0f), ost = st;
    for (int blk = 0; blk < 2; ++blk) {
      std::vector<float> x = randv(5600), y, wy(1029);
      convolveBlockResampleFIR(y, x, h, st, 800, 147, false);
      orc_fir_resample(wy.data(), x.data(), x.size(), h.data(), h.size(), ost.data(), 800, 147);
      expect_same("convolveBlockResampleFIR", y, wy);
      expect_same("resampler state", st, ost);
    }
  }
  {
    // a state vector the C ABI cannot honour (it assumes h.size() - 1 floats) is refused, not overrun
    std::vector<float> x = randv(64), y, bad_state(7, 0.0f);
    bool threw = false;
    try {
      convolveBlockFIR(y, x, pilot, bad_state);
    } catch (const std::invalid_argument &) {
      threw = true;
    }
    std::printf("%-34s %s\n", "state of the wrong size refused", threw ? "ok" : "NOT REFUSED");
    if (!threw) ++failures;
  }
  std::printf("%s\n", failures ? "FAILED" : "ALL OK");
  return failures ? 1 : 0;
}
