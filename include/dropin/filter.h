/*
 * include/dropin/filter.h -- source-compatible replacement for the reference's
 * include/filter.h (lines 18-43).  A translation unit that includes this header
 * instead of the reference's and links libsdr_filter.so + libsdr_b200.so runs every
 * filter.h call on the B200 with bit-identical results; nothing else in the caller
 * changes (same names, same argument order and meaning, same in/out conventions:
 * outputs are cleared and resized by the callee, `state` vectors are caller-owned,
 * updated in place, and their size encodes the tap count).
 *
 * Differences, all deliberate:
 *   - a CUDA failure (no sm_100 device, out of memory) throws std::runtime_error
 *     carrying sdr_last_error(); the reference has no failure mode to mirror and
 *     there is no CPU fallback;
 *   - convolveBlockFastFIR does not perform the reference's one-past-the-end
 *     iteration (src/filter.cpp:166), so it neither reads x[x.size()] nor writes
 *     y[x.size()/decim]; every in-range output is identical;
 *   - a `state` vector whose size is not h.size()-1 (6 for fmPLL) throws
 *     std::invalid_argument: the reference would index it by state.size() and treat
 *     that as the tap count (filter.cpp:144,174,207); the C ABI below takes a bare pointer;
 *   - the device is picked by the SDR_B200_DEVICE environment variable (default 0).
 */
#ifndef SDR_B200_DROPIN_FILTER_H
#define SDR_B200_DROPIN_FILTER_H

#include <cmath>
#include <iostream>
#include <vector>

/* ---- coefficient design (host) ------------------------------------------- */
/* reference filter.h:24 / filter.cpp:103-114 */
void impulseResponseLPF(float Fs, float Fc, unsigned short int num_taps, std::vector<float> &h);
/* reference filter.h:20 / filter.cpp:83-99 */
void bandPass(float Fs, float Fb, float Fe, unsigned short int N_taps, std::vector<float> &coeff);

/* ---- convolutions --------------------------------------------------------- */
/* reference filter.h:26 / filter.cpp:118-130 */
void convolveFIR(std::vector<float> &y, const std::vector<float> &x, const std::vector<float> &h);
/* reference filter.h:28 / filter.cpp:133-154 */
void convolveBlockFIR(std::vector<float> &y, const std::vector<float> &x,
                      const std::vector<float> &h, std::vector<float> &state);
/* reference filter.h:31 / filter.cpp:158-188 (printData is ignored there too) */
void convolveBlockFastFIR(std::vector<float> &y, const std::vector<float> &x,
                          const std::vector<float> &h, std::vector<float> &state,
                          const unsigned int decim, const bool printData);
/* reference filter.h:34 / filter.cpp:191-223 */
void convolveBlockResampleFIR(std::vector<float> &y, const std::vector<float> &x,
                              const std::vector<float> &h, std::vector<float> &state,
                              const unsigned int audio_decim, const unsigned int audio_upsamp,
                              bool printData);
/* reference filter.h:37 / filter.cpp:227-234 */
void upsample(const std::vector<float> &x, std::vector<float> &xu, const int up_rate);
/* reference filter.h:39 / filter.cpp:237-245 */
void downsample(std::vector<float> &output, const std::vector<float> &input,
                const unsigned short int ds_coeff);

/* ---- demodulation / carrier recovery -------------------------------------- */
/* reference filter.h:41 / filter.cpp:248-266 */
void fmDemod(std::vector<float> &fm_demod, const std::vector<float> &I, const std::vector<float> &Q,
             float &prev_i, float &prev_q);
/* reference filter.h:22 / filter.cpp:32-80 */
void fmPLL(const std::vector<float> &PLLIn, std::vector<float> &ncoOut, std::vector<float> &state,
           float freq, float Fs, float ncoScale, float phaseAdjust, float normBandwidth);
/* reference filter.h:18 / filter.cpp:14-29 */
void allPass(const std::vector<float> &input_block, std::vector<float> &state_block,
             std::vector<float> &output_block);

/* ---- helper ---------------------------------------------------------------- */
/* reference filter.h:43 declares `int mode=1`, filter.cpp:270 defines `unsigned short`;
 * this is the declared form, with the defined behaviour (host-side range copy). */
void setVec(const std::vector<float> &vec1, std::vector<float> &vec2, int begin, int end,
            int mode = 1);

#endif /* SDR_B200_DROPIN_FILTER_H */
