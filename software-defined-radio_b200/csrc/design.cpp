// design.cpp -- FIR coefficient design on the host (stays C++ like the reference;
// it runs once per pipeline and is not part of the data path).
//
// The reference evaluates these expressions in double (its PI is a double macro,
// include/dy4.h:23) and rounds once per store into the float vector; the window
// is sin^2(i*pi/N).  Compile with -ffp-contract=off so that no multiply-add is
// fused: the reference is built for baseline x86-64 where none can be.
#include "design.h"

#include <cmath>

#include "../../include/sdr_b200.h"

namespace sdr {

static const double kPi = 3.14159265358979323846;

// Windowed-sinc low-pass with cut-off Fc; the centre tap index is (ntaps-1)/2 in
// integer arithmetic, so even tap counts are not symmetric (as in the reference).
void design_lpf(float Fs, float Fc, unsigned short ntaps, float *h) {
  const float cutoff = Fc / (Fs / 2);  // normalised to Nyquist, in float
  const int mid = (ntaps - 1) / 2;
  for (int i = 0; i < ntaps; ++i) {
    float tap = cutoff;
    if (i != mid) {
      const double a = kPi * cutoff * (i - mid);
      tap = (float)(cutoff * (std::sin(a) / a));
    }
    const double win = std::sin(i * kPi / ntaps);
    h[i] = (float)(tap * (win * win));
  }
}

// Band-pass between Fb and Fe: a low-pass of half the pass width shifted to the
// band centre by a cosine, same window.  Note the three separate float stores.
void design_bpf(float Fs, float Fb, float Fe, unsigned short ntaps, float *h) {
  const float centre = ((Fe + Fb) / 2) / (Fs / 2);
  const float width = (Fe - Fb) / (Fs / 2);
  const int mid = (ntaps - 1) / 2;
  for (int i = 0; i < ntaps; ++i) {
    float tap = width;
    if (i != mid) {
      const double a = kPi * width / 2 * (i - mid);
      tap = (float)(width * (std::sin(a) / a));
    }
    tap = (float)(tap * std::cos(i * kPi * centre));
    const double win = std::sin(i * kPi / ntaps);
    h[i] = (float)(tap * win * win);
  }
}

}  // namespace sdr

extern "C" int sdr_lpf_design(float Fs, float Fc, unsigned short ntaps, float *h) {
  if (!h || ntaps == 0) return SDR_ERR_INVALID;
  sdr::design_lpf(Fs, Fc, ntaps, h);
  return SDR_OK;
}

extern "C" int sdr_bpf_design(float Fs, float Fb, float Fe, unsigned short ntaps, float *h) {
  if (!h || ntaps == 0) return SDR_ERR_INVALID;
  sdr::design_bpf(Fs, Fb, Fe, ntaps, h);
  return SDR_OK;
}
