/*
 * fm_oracle.h -- CPU restatement of the reference FM receiver DSP chain.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and there only as the checker or the timed CPU baseline.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_*.py)
 * bit-for-bit against the reference's own sources compiled unmodified into
 * oracle/_ref/libfmref.so (see oracle/Makefile) and against the golden vectors
 * under tests/golden/ that were generated from that library.
 *
 * All file:line citations are relative to the upstream reference tree.
 */
#ifndef FM_ORACLE_H
#define FM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- libm replicas (glibc 2.39, x86-64) used by fmPLL ------------------- */
/* atan2f: sysdeps/ieee754/flt-32/e_atan2f.c + s_atanf.c (float, no FMA).   */
float orc_atan2f(float y, float x);
/* sincosf / cosf: sysdeps/ieee754/flt-32/s_sincosf.c, FMA ifunc variant.    */
void orc_sincosf(float x, float *s, float *c);
float orc_cosf(float x);
void orc_atan2f_batch(const float *y, const float *x, size_t n, float *out);
void orc_sincosf_batch(const float *x, size_t n, float *s, float *c);

/* ---- filter design (src/filter.cpp:83-114) ------------------------------ */
void orc_lpf_design(float Fs, float Fc, unsigned short ntaps, float *h);
void orc_bpf_design(float Fs, float Fb, float Fe, unsigned short ntaps, float *h);

/* ---- primitives (src/filter.cpp) ---------------------------------------- */
/* iofunc.cpp:128-135 */
void orc_u8_to_f32(const uint8_t *raw, size_t n, float *out);
/* filter.cpp:133-154.  state has nh-1 entries. y has nx entries. */
void orc_fir_block(float *y, const float *x, size_t nx, const float *h, size_t nh,
                   float *state);
/* filter.cpp:158-188 without the one-past-the-end iteration. y has nx/decim. */
void orc_fir_decim(float *y, const float *x, size_t nx, const float *h, size_t nh,
                   float *state, unsigned decim);
/* filter.cpp:191-223.  state is the reference's zero-stuffed nh-1 vector. */
void orc_fir_resample(float *y, const float *x, size_t nx, const float *h, size_t nh,
                      float *state, unsigned decim, unsigned upsamp);
/* filter.cpp:248-266 */
void orc_fm_demod(float *out, const float *I, const float *Q, size_t n,
                  float *prev_i, float *prev_q);
/* filter.cpp:14-29.  state has ns entries. out has n entries. */
void orc_allpass(const float *in, size_t n, float *state, size_t ns, float *out);
/* filter.cpp:32-80. out has n+1 entries; state has 6. */
void orc_pll(const float *in, size_t n, float *out, float *state, float freq, float Fs,
             float ncoScale, float phaseAdjust, float normBandwidth);
/* threadMonoOnly.cpp:185-190 (x86-64 cvttss2si semantics) */
int16_t orc_pcm16(float v);

/* ---- whole chain (project.cpp:40-152,154-309,311-382,421-458) ----------- */
typedef struct {
  int mode;        /* 0..3 */
  int channels;    /* 1 mono, 2 stereo */
  int rf_taps;     /* 151 functional / 13 as shipped */
  int audio_taps;  /* per-phase count: 101 functional / 13 as shipped */
  int stereo_taps; /* 151 functional / 13 as shipped */
} orc_config;

typedef struct {
  int rf_Fs, if_Fs, audio_Fs;
  int rf_decim, audio_decim, audio_upsamp; /* audio_upsamp==1 for modes 0/1 */
  int block_bytes;                         /* reference block size in bytes */
} orc_mode_info;

int orc_mode_lookup(int mode, orc_mode_info *out);

typedef struct orc_chain orc_chain;

/* Stage ids for orc_chain_tap / sdr_pipeline_tap (same numbering both sides). */
enum {
  ORC_TAP_I_FILT = 0,
  ORC_TAP_Q_FILT = 1,
  ORC_TAP_DEMOD = 2,
  ORC_TAP_ALLPASS = 3,
  ORC_TAP_STEREO_FILT = 4,
  ORC_TAP_CARRIER_FILT = 5,
  ORC_TAP_NCO = 6,
  ORC_TAP_MIXER = 7,
  ORC_TAP_AUDIO_FILT = 8,
  ORC_TAP_STEREO_FINAL = 9,
  ORC_TAP_COUNT = 10
};

orc_chain *orc_chain_create(const orc_config *cfg);
void orc_chain_destroy(orc_chain *c);
void orc_chain_reset(orc_chain *c);
/* Processes floor(nbytes/block_bytes) reference blocks, block by block exactly
 * as the reference does.  pcm receives mono samples or L,R interleaved; returns
 * the number of int16 values written.  If keep_taps != 0 the float intermediates
 * of every block are appended to growable per-stage buffers. */
size_t orc_chain_process(orc_chain *c, const uint8_t *iq, size_t nbytes, int16_t *pcm,
                         int keep_taps);
/* Returns pointer/length of the accumulated intermediate for a stage. */
const float *orc_chain_tap(const orc_chain *c, int stage, size_t *n);
void orc_chain_clear_taps(orc_chain *c);

#ifdef __cplusplus
}
#endif
#endif
