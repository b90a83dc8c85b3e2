/*
 * rds_oracle.h -- CPU restatement (double precision) of the reference's RDS model.
 *
 * TEST INFRASTRUCTURE ONLY (see fm_oracle.h): only tests/, __graft_entry__.smoke() and
 * bench.py's CPU legs may load it, as the checker or the timed CPU baseline.
 *
 * The reference has no C++ RDS path; its RDS receiver is the Python model
 * model/fmRDS.py:222-278 on top of model/fmSupportLib.py.
 * This file restates those functions in plain C, each citing the lines it follows.
 *
 * Parity status: PINNED against the reference itself: tests/golden/make_golden_rds.py
 * imports /root/reference/model/fmSupportLib.py (numpy + scipy.signal.lfilter, as
 * fmRDS.py does) in the build container and stores its outputs in tests/golden/rds_*.npz;
 * tests/test_oracle_rds.py checks every function below against them (floating-point
 * stages to 1e-12 of full scale -- numpy/scipy sum in another order -- bit layer exactly),
 * plus the five offset-word syndromes of the RDS standard (fmSupportLib.py:32-57).
 *
 * All file:line citations are relative to the upstream reference tree.
 */
#ifndef RDS_ORACLE_H
#define RDS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- filter design (double) ---------------------------------------------- */
/* fmSupportLib.py:358-372 */
void rdo_bandpass(int ntaps, double Fs, double Fb, double Fe, double *h);
/* fmSupportLib.py:376-385 (also model/fmMonoBasic.py impResponse, the one fmRDS.py imports) */
void rdo_lowpass(int ntaps, double Fs, double Fc, double *h);
/* fmSupportLib.py:251-287 */
void rdo_rrc(double Fs, int ntaps, double *h);

/* ---- streaming primitives -------------------------------------------------- */
/* scipy.signal.lfilter(h, 1.0, x, zi=state) as used at fmRDS.py:223,233,248: a direct-form
 * FIR whose carried state is equivalent to the last nh-1 inputs.  hist has nh-1 entries
 * (oldest first), updated in place. */
void rdo_fir(const double *x, size_t n, const double *h, int nh, double *hist, double *y);
/* fmSupportLib.py:291-295.  state has ns entries. */
void rdo_allpass(const double *x, size_t n, double *state, int ns, double *y);
/* fmSupportLib.py:297-354.  state has 7 entries {integrator, phaseEst, feedbackI, feedbackQ,
 * ncoOut[-1], trigOffset, ncoOutQ[-1]}; outI/outQ have n+1 entries. */
void rdo_pll(const double *x, size_t n, double freq, double Fs, double *state, double ncoScale,
             double phaseAdjust, double normBandwidth, double *outI, double *outQ);
/* fmSupportLib.py:388-407 (gain U, not 1+U as in src/filter.cpp:213).  hist holds the last
 * nh/U - 1 ... see the .c file: it is the compact form of the zero-stuffed state. */
void rdo_resample(const double *x, size_t n, const double *h, int nh, double *hist, int decim,
                  int upsamp, double *y);

/* ---- bit layer --------------------------------------------------------------- */
/* fmSupportLib.py:103-201 with the state the driver passes (fmRDS.py:257-260: pair = 0,0;
 * start = 158; prev_size = 0, re-initialised for every block).  Returns the number of bits
 * written to bits (capacity cap).  Where the reference would loop forever (no pair with
 * opposite signs in a whole pass) the pass is accepted as it stands. */
int rdo_cdr(const double *x, int n, int sps, int block_count, uint8_t *bits, int cap);
/* The same function with its to_pass_on_state explicit and carried: state = {pair[0], pair[1],
 * start, prev_size} (fmSupportLib.py:104-106), updated as at :178-189. */
int rdo_cdr_state(const double *x, int n, int sps, int block_count, double *state, uint8_t *bits, int cap);
/* fmSupportLib.py:241-249 */
void rdo_diff_decode(const uint8_t *in, int n, uint8_t *out);
/* fmSupportLib.py:14-27 with the parity matrix of :32-57; d has 26 bits, s gets 10. */
void rdo_syndrome(const uint8_t *d, uint8_t *s);
/* fmSupportLib.py:30-100.  Returns the offset type as a character (' ', 'A', 'B', 'C',
 * 'c' for C', 'D'); *state_index is the number of bits consumed. */
char rdo_framesync(const uint8_t *d, int n, int *state_index);

/* ---- whole chain from fm_demod (fmRDS.py:222-274) ------------------------------ */
typedef struct rdo_chain rdo_chain;
/* mode 0 or 2 (fmRDS.py:55-75).  block_if = IF samples per block (fmRDS.py:149-152 divided
 * by 2*rf_decim): 9600 in mode 0, 1536000 in mode 2; any multiple of 960 / 1920 works. */
rdo_chain *rdo_chain_create(int mode, int block_if);
void rdo_chain_destroy(rdo_chain *c);
/* Keep the CDR state from block to block (what fmSupportLib.py's CDR was written for) instead of
 * re-creating it per block as fmRDS.py:257-260 does.  Off by default. */
void rdo_chain_set_cdr_carry(rdo_chain *c, int carry);
/* One block of fm_demod in; every intermediate is kept until the next call. */
void rdo_chain_block(rdo_chain *c, const double *fm_demod);
/* stage: 0 channel_filt, 1 carrier_filt, 2 PLL I (n+1), 3 PLL Q (n+1), 4 mixer I, 5 mixer Q,
 * 6 resampler I, 7 resampler Q, 8 RRC I, 9 RRC Q.  Returns a pointer and the length. */
const double *rdo_chain_tap(const rdo_chain *c, int stage, size_t *n);
/* Bits of the last block: CDR + Manchester output, then the differential decoding of it. */
int rdo_chain_bits(const rdo_chain *c, const uint8_t **cdr_bits, const uint8_t **diff_bits);
/* Frame synchroniser result for the last block (fmRDS.py:270-274). */
char rdo_chain_offset(const rdo_chain *c);

#ifdef __cplusplus
}
#endif
#endif
