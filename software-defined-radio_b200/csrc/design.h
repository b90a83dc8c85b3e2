// design.h -- host-side filter design shared by the pipeline and the C ABI.
#pragma once
namespace sdr {
// impulseResponseLPF, src/filter.cpp:103-114
void design_lpf(float Fs, float Fc, unsigned short ntaps, float *h);
// bandPass, src/filter.cpp:83-99
void design_bpf(float Fs, float Fb, float Fe, unsigned short ntaps, float *h);
}  // namespace sdr
