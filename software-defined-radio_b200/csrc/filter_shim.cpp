// filter_shim.cpp -- the reference's C++ vector API (include/dropin/filter.h) as thin
// wrappers over the C ABI (include/sdr_b200.h).  Each wrapper only adapts the calling
// convention: resize the out-vector as the reference does, pass raw pointers down,
// surface failures as exceptions.
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "dropin/filter.h"
#include "sdr_b200.h"

namespace {

int device() {
  static const int d = [] {
    const char *e = std::getenv("SDR_B200_DEVICE");
    return e ? std::atoi(e) : 0;
  }();
  return d;
}

void must(int rc, const char *what) {
  if (rc != SDR_OK) throw std::runtime_error(std::string(what) + ": " + sdr_last_error());
}

// Vectors may be empty; the ABI wants valid pointers only when sizes are non-zero.
template <typename V>
auto ptr(V &v) -> decltype(v.data()) {
  static typename V::value_type dummy[4];
  return v.empty() ? dummy : v.data();
}

// The C ABI takes the state as a bare pointer whose length is implied by the tap count
// (nh-1 floats; the zero-stuffed nh-1 for the resampler; 6 for the PLL).  The reference
// indexes by state.size() instead (filter.cpp:144,174,207), so a vector of another size would
// be read and written out of bounds below the ABI: refuse it here.
void need_state(const std::vector<float> &state, size_t want, const char *what) {
  if (state.size() != want)
    throw std::invalid_argument(std::string(what) + ": state must hold " + std::to_string(want) +
                                " floats (got " + std::to_string(state.size()) + ")");
}

}  // namespace

void impulseResponseLPF(float Fs, float Fc, unsigned short int num_taps, std::vector<float> &h) {
  h.assign(num_taps, 0.0f);
  if (num_taps) must(sdr_lpf_design(Fs, Fc, num_taps, h.data()), "impulseResponseLPF");
}

void bandPass(float Fs, float Fb, float Fe, unsigned short int N_taps, std::vector<float> &coeff) {
  coeff.assign(N_taps, 0.0f);
  if (N_taps) must(sdr_bpf_design(Fs, Fb, Fe, N_taps, coeff.data()), "bandPass");
}

void convolveFIR(std::vector<float> &y, const std::vector<float> &x, const std::vector<float> &h) {
  y.assign(x.size() + h.size() - 1, 0.0f);
  if (x.empty() || h.empty()) return;
  must(sdr_convolve(device(), y.data(), x.data(), x.size(), h.data(), h.size()), "convolveFIR");
}

void convolveBlockFIR(std::vector<float> &y, const std::vector<float> &x,
                      const std::vector<float> &h, std::vector<float> &state) {
  y.assign(x.size(), 0.0f);
  if (x.empty()) return;
  need_state(state, h.empty() ? 0 : h.size() - 1, "convolveBlockFIR");
  must(sdr_fir_block(device(), y.data(), x.data(), x.size(), h.data(), h.size(), ptr(state)),
       "convolveBlockFIR");
}

void convolveBlockFastFIR(std::vector<float> &y, const std::vector<float> &x,
                          const std::vector<float> &h, std::vector<float> &state,
                          const unsigned int decim, const bool) {
  y.assign(x.size() / decim, 0.0f);
  if (x.empty()) return;
  need_state(state, h.empty() ? 0 : h.size() - 1, "convolveBlockFastFIR");
  must(sdr_fir_decim(device(), ptr(y), x.data(), x.size(), h.data(), h.size(), ptr(state), decim),
       "convolveBlockFastFIR");
}

void convolveBlockResampleFIR(std::vector<float> &y, const std::vector<float> &x,
                              const std::vector<float> &h, std::vector<float> &state,
                              const unsigned int audio_decim, const unsigned int audio_upsamp, bool) {
  y.assign((x.size() * audio_upsamp) / audio_decim, 0.0f);
  if (x.empty()) return;
  need_state(state, h.empty() ? 0 : h.size() - 1, "convolveBlockResampleFIR");
  must(sdr_fir_resample(device(), ptr(y), x.data(), x.size(), h.data(), h.size(), ptr(state),
                        audio_decim, audio_upsamp),
       "convolveBlockResampleFIR");
}

void upsample(const std::vector<float> &x, std::vector<float> &xu, const int up_rate) {
  xu.assign(x.size() * up_rate, 0.0f);
  if (x.empty()) return;
  must(sdr_upsample(device(), x.data(), x.size(), xu.data(), up_rate), "upsample");
}

void downsample(std::vector<float> &output, const std::vector<float> &input,
                const unsigned short int ds_coeff) {
  output.assign((size_t)std::ceil(input.size() / static_cast<float>(ds_coeff)), 0.0f);
  if (input.empty()) return;
  must(sdr_downsample(device(), output.data(), input.data(), input.size(), ds_coeff), "downsample");
}

void fmDemod(std::vector<float> &fm_demod, const std::vector<float> &I, const std::vector<float> &Q,
             float &prev_i, float &prev_q) {
  fm_demod.assign(I.size(), 0.0f);
  if (I.empty()) return;
  must(sdr_fm_demod(device(), fm_demod.data(), I.data(), Q.data(), I.size(), &prev_i, &prev_q),
       "fmDemod");
}

void fmPLL(const std::vector<float> &PLLIn, std::vector<float> &ncoOut, std::vector<float> &state,
           float freq, float Fs, float ncoScale, float phaseAdjust, float normBandwidth) {
  ncoOut.assign(PLLIn.size() + 1, 0.0f);
  need_state(state, 6, "fmPLL");
  must(sdr_pll(device(), ptr(PLLIn), PLLIn.size(), ncoOut.data(), state.data(), freq, Fs, ncoScale,
               phaseAdjust, normBandwidth),
       "fmPLL");
}

void allPass(const std::vector<float> &input_block, std::vector<float> &state_block,
             std::vector<float> &output_block) {
  output_block.assign(input_block.size(), 0.0f);
  if (input_block.empty()) return;
  must(sdr_allpass(device(), input_block.data(), input_block.size(), ptr(state_block),
                   state_block.size(), output_block.data()),
       "allPass");
}

// Pure host bookkeeping in the reference as well (never called from project.cpp).
void setVec(const std::vector<float> &vec1, std::vector<float> &vec2, int begin, int end, int mode) {
  if (mode == 1) {
    for (int i = begin, k = 0; i < end; ++i, ++k) vec2[k] = vec1[i];
  } else if (mode == 2) {
    for (int i = begin, k = 0; i < end; ++i, ++k) vec2[i] = vec1[k];
  }
}
