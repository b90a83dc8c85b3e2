/*
 * sdr_b200.h -- C ABI of the B200-native FM receiver DSP path.
 *
 * This is the drop-in boundary.  The reference (mnigm2001/Software-Defined-Radio)
 * has no FFI layer; its "operator API" for this path is the set of C++ free
 * functions in include/filter.h:18-43 plus the process contract of
 * src/project.cpp:385-500 (raw interleaved uint8 I/Q on stdin, native-endian
 * int16 PCM on stdout, mode 0-3, 1|2 audio channels).  Every entry point below
 * names the reference interface it replaces.  include/dropin/filter.h re-declares
 * the reference's C++ prototypes on top of this ABI.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross the boundary;
 *   - every function returns an int status: 0 (SDR_OK) or a negative SDR_ERR_*;
 *     sdr_last_error() returns a thread-local message for the last failure;
 *   - the caller owns every host buffer; the library owns device state;
 *   - one sdr_pipeline handle must be driven by one thread at a time; distinct
 *     handles are independent;
 *   - there is NO CPU fallback: every compute entry point fails with
 *     SDR_ERR_NO_DEVICE when no sm_100 CUDA device is usable.
 *
 * All arithmetic that the reference performs in float is performed on the device
 * in the same order with the same roundings (separate multiply and add, IEEE
 * division, glibc-identical atan2f/sincosf/cosf), so results are bit-identical
 * to the reference built with g++ -O3 on x86-64/glibc 2.39.
 */
#ifndef SDR_B200_H
#define SDR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDR_OK 0
#define SDR_ERR_INVALID -1   /* bad argument */
#define SDR_ERR_NO_DEVICE -2 /* no usable sm_100 device / CUDA driver */
#define SDR_ERR_CUDA -3      /* a CUDA call failed; see sdr_last_error() */
#define SDR_ERR_NOMEM -4     /* host or device allocation failed */
#define SDR_ERR_CAPACITY -5  /* request exceeds the capacity fixed at create time */

/* Arithmetic variants of the batched pipeline. */
#define SDR_VARIANT_EXACT 0 /* CUDA-core path, bit-identical to the reference */
/* Tensor-core RF front end (tcgen05 kind::i8 over the raw byte stream), mono only, all four
 * modes (rf_decim 10, 5, 3), rf_taps <= 151.  I/Q are fixed-point FIR outputs (tap
 * quantisation 2^-34, exact integer sums, <= 1 ulp recombination) instead of the reference's
 * sequential float sums, and the audio filter contracts its multiply-adds: float
 * intermediates agree to >= 100 dB SNR and PCM to +-1 LSB, not bit for bit.
 * Stereo is refused: the PLL amplifies 1-ulp differences beyond the parity bound. */
#define SDR_VARIANT_FAST 1
/* Mono or stereo.  Everything that feeds the stereo PLL -- RF front end, discriminator, pilot
 * band-pass, the PLL itself -- keeps the reference's arithmetic bit for bit (so the NCO is the
 * reference's NCO); the filters that do NOT feed it (22-54 kHz band-pass, both audio low-pass /
 * resampling filters) contract each multiply-add into one FMA.  fm_demod, carrier_filt and the
 * NCO are bit-identical, the other float intermediates agree to >= 100 dB and PCM to +-1 LSB. */
#define SDR_VARIANT_MIXED 2
/* Flag, OR-ed into a variant: run the exact 151-tap FIR kernels (RF front end, stereo band-pass
 * pair) in their scalar form (FMUL + FADD per tap) instead of the packed one (FMUL2 + FFMA2 on
 * register pairs).  Same bits either way; kept so that the two forms can be checked against each
 * other and timed side by side. */
#define SDR_VARIANT_SCALAR_FIR 0x100

/* Intermediate signals that sdr_pipeline_tap can return (same numbering as the
 * oracle, oracle/fm_oracle.h).  Names follow src/project.cpp's variables. */
#define SDR_TAP_I_FILT 0        /* project.cpp:111  */
#define SDR_TAP_Q_FILT 1        /* project.cpp:121  */
#define SDR_TAP_DEMOD 2         /* project.cpp:128  fm_demod */
#define SDR_TAP_ALLPASS 3       /* project.cpp:194  audio_allpass */
#define SDR_TAP_STEREO_FILT 4   /* project.cpp:202  stereo_filt */
#define SDR_TAP_CARRIER_FILT 5  /* project.cpp:207  carrier_filt */
#define SDR_TAP_NCO 6           /* project.cpp:237  PLL[0..N) as seen by the mixer */
#define SDR_TAP_MIXER 7         /* project.cpp:246-248 */
#define SDR_TAP_AUDIO_FILT 8    /* project.cpp:219/227/346/353 audio_filt */
#define SDR_TAP_STEREO_FINAL 9  /* project.cpp:257/264 stereo_final */
#define SDR_TAP_COUNT 10

const char *sdr_version(void);
const char *sdr_last_error(void);
/* Number of usable sm_100 devices (0 when there is none; never fails). */
int sdr_device_count(void);

/* ------------------------------------------------------------------------ */
/* Filter design (host; double precision internally, like the reference).   */
/* ------------------------------------------------------------------------ */
/* Replaces impulseResponseLPF(Fs, Fc, num_taps, h)   include/filter.h:24, src/filter.cpp:103-114 */
int sdr_lpf_design(float Fs, float Fc, unsigned short ntaps, float *h);
/* Replaces bandPass(Fs, Fb, Fe, N_taps, coeff)        include/filter.h:20, src/filter.cpp:83-99  */
int sdr_bpf_design(float Fs, float Fb, float Fe, unsigned short ntaps, float *h);

/* ------------------------------------------------------------------------ */
/* Single stateful operators on HOST buffers (upload -> kernel -> download). */
/* They exist so that the reference's filter.h functions can be re-pointed   */
/* one by one; the batched pipeline below is the throughput path.            */
/* `device` is a CUDA ordinal.  State buffers are updated in place with the  */
/* reference's layouts so callers may mix these with the reference freely.   */
/* ------------------------------------------------------------------------ */
/* convolveFIR(y, x, h)                                filter.h:26, filter.cpp:118-130; y has nx+nh-1 */
int sdr_convolve(int device, float *y, const float *x, size_t nx, const float *h, size_t nh);
/* convolveBlockFIR(y, x, h, state)                    filter.h:28, filter.cpp:133-154; state nh-1, y nx */
int sdr_fir_block(int device, float *y, const float *x, size_t nx, const float *h, size_t nh,
                  float *state);
/* convolveBlockFastFIR(y, x, h, state, decim, _)      filter.h:31, filter.cpp:158-188; y nx/decim.
 * The reference's one-past-the-end iteration (filter.cpp:166) is not performed. */
int sdr_fir_decim(int device, float *y, const float *x, size_t nx, const float *h, size_t nh,
                  float *state, unsigned decim);
/* convolveBlockResampleFIR(y, x, h, state, D, U, _)   filter.h:34, filter.cpp:191-223;
 * state is the reference's zero-stuffed nh-1 vector; y has nx*U/D; gain is (1+U). */
int sdr_fir_resample(int device, float *y, const float *x, size_t nx, const float *h, size_t nh,
                     float *state, unsigned decim, unsigned upsamp);
/* fmDemod(out, I, Q, prev_i, prev_q)                  filter.h:41, filter.cpp:248-266 */
int sdr_fm_demod(int device, float *out, const float *I, const float *Q, size_t n, float *prev_i,
                 float *prev_q);
/* fmPLL(in, ncoOut, state, freq, Fs, ncoScale, phaseAdjust, normBandwidth)
 *                                                     filter.h:22, filter.cpp:32-80; out n+1, state 6 */
int sdr_pll(int device, const float *in, size_t n, float *out, float *state, float freq, float Fs,
            float ncoScale, float phaseAdjust, float normBandwidth);
/* allPass(in, state, out)                             filter.h:18, filter.cpp:14-29; state ns <= n */
int sdr_allpass(int device, const float *in, size_t n, float *state, size_t ns, float *out);
/* upsample(x, xu, U)                                  filter.h:37, filter.cpp:227-234; xu has nx*U */
int sdr_upsample(int device, const float *x, size_t nx, float *xu, int up_rate);
/* downsample(out, in, D)                              filter.h:39, filter.cpp:237-245; out ceil(n/D) */
int sdr_downsample(int device, float *out, const float *in, size_t n, unsigned short ds);

/* ------------------------------------------------------------------------ */
/* Batched receiver pipeline: project.cpp's RF_FrontEnd + RF_MONO/RF_STEREO  */
/* for `batch` independent captures, all state carried on the device.        */
/* ------------------------------------------------------------------------ */
typedef struct sdr_pipeline sdr_pipeline;

typedef struct {
  int mode;        /* 0..3: project.cpp:424-427 mode table */
  int channels;    /* 1 mono (RF_MONO), 2 stereo (RF_STEREO) */
  int rf_taps;     /* project.cpp:46 (13 as shipped) / threadMonoOnly.cpp:66 (151) */
  int audio_taps;  /* per-phase count; x audio_upsamp in modes 2/3 (project.cpp:425-426) */
  int stereo_taps; /* project.cpp:429 */
  int batch;       /* independent captures processed per call (>= 1) */
  int device;      /* CUDA ordinal */
  int variant;     /* SDR_VARIANT_* */
  uint64_t max_bytes_per_channel; /* capacity of one process call, per capture */
} sdr_config;

typedef struct {
  int rf_Fs, if_Fs, audio_Fs;
  int rf_decim, audio_decim, audio_upsamp; /* audio_upsamp is 1 in modes 0/1 */
  int block_bytes;   /* the reference's block size (project.cpp:55-57) */
  int granule_bytes; /* process() accepts any multiple of this per capture */
  int pcm_per_granule; /* int16 values produced per granule (x2 when stereo) */
} sdr_mode_info;

int sdr_mode_lookup(int mode, int channels, sdr_mode_info *out);

int sdr_pipeline_create(const sdr_config *cfg, sdr_pipeline **out);
int sdr_pipeline_destroy(sdr_pipeline *p);
/* Back to the reference's initial state (project.cpp:61-65,446-458). */
int sdr_pipeline_reset(sdr_pipeline *p);
/* Copies every carried state of capture `src` over capture `dst` (used to
 * restart / fork a capture; no reference equivalent). */
int sdr_pipeline_copy_state(sdr_pipeline *p, int dst, int src);

/* Number of int16 values produced per capture for nbytes_per_channel input. */
int sdr_pipeline_pcm_count(const sdr_pipeline *p, size_t nbytes_per_channel, size_t *n_pcm);

/* Device-resident form.  d_iq is [batch][iq_stride_bytes] uint8 interleaved I,Q
 * in device memory (only the first nbytes_per_channel of each row are read);
 * d_pcm is [batch][pcm_stride] int16 (mono) or L,R interleaved (stereo).
 * `stream` is a cudaStream_t (NULL = default stream); the call only enqueues. */
int sdr_pipeline_process_device(sdr_pipeline *p, const uint8_t *d_iq, size_t iq_stride_bytes,
                                size_t nbytes_per_channel, int16_t *d_pcm, size_t pcm_stride,
                                void *stream);

/* Host form (the call a user of the reference's `project` would make for a
 * batch): uploads in slices over pinned double buffers on two streams, runs the
 * pipeline, downloads PCM, and returns when pcm is complete. */
int sdr_pipeline_process_host(sdr_pipeline *p, const uint8_t *iq, size_t iq_stride_bytes,
                              size_t nbytes_per_channel, int16_t *pcm, size_t pcm_stride);

/* Keep (1) or drop (0, default) the float intermediates of the NEXT process
 * calls so that sdr_pipeline_tap can return them (costs extra HBM traffic). */
int sdr_pipeline_keep_taps(sdr_pipeline *p, int keep);
/* Copies intermediate `stage` of capture `channel` from the last process call
 * to host memory; *n receives the element count (dst may be NULL to query). */
int sdr_pipeline_tap(sdr_pipeline *p, int stage, int channel, float *dst, size_t cap, size_t *n);

/* Number of kernels this handle has launched since creation / last reset of the
 * counter (bench.py reports it as gpu_launches). */
int sdr_pipeline_launch_count(sdr_pipeline *p, uint64_t *count, int reset);

/* Page-locked host memory for callers that have no CUDA headers (the C++ CLI):
 * buffers from sdr_host_alloc are DMA-able, so process_host skips its staging copy. */
int sdr_host_alloc(size_t bytes, void **out);
int sdr_host_free(void *ptr);

/* Per-kernel device timing (CUDA events on the launching stream, recorded around
 * every kernel of this handle while enabled).  sdr_pipeline_kernel_times walks the
 * kernels by index: returns 0 and fills name/total/count, or 1 past the last one. */
int sdr_pipeline_profile(sdr_pipeline *p, int enable);
int sdr_pipeline_kernel_times(sdr_pipeline *p, int index, char *name, size_t name_cap,
                              double *total_ms, uint64_t *count, int reset);

/* ------------------------------------------------------------------------ */
/* One process, several devices.  Replaces the thread set-up of              */
/* src/project.cpp:471-496 (one producer + one consumer around a queue) at    */
/* the scale of a batch: `cfg.batch` captures are cut into contiguous ranges, */
/* one per device, each driven by its own host thread through                 */
/* sdr_pipeline_process_host.  Captures never interact, so there is no        */
/* collective; every device writes its PCM rows straight into the caller's    */
/* buffer (the final host gather).  Results are identical to one pipeline     */
/* processing the whole batch.  cfg.device is ignored.                        */
/* ------------------------------------------------------------------------ */
typedef struct sdr_multi sdr_multi;
typedef struct {
  sdr_config cfg;     /* cfg.batch = captures over all devices */
  int n_devices;      /* 0 = every usable sm_100 device; clipped to cfg.batch */
  const int *devices; /* optional CUDA ordinals (n_devices entries); NULL = 0..n_devices-1 */
} sdr_multi_config;

int sdr_multi_create(const sdr_multi_config *cfg, sdr_multi **out);
int sdr_multi_destroy(sdr_multi *m);
int sdr_multi_reset(sdr_multi *m);
/* Devices in use and the first capture of each device's range (up to `cap` entries each). */
int sdr_multi_layout(const sdr_multi *m, int *n_devices, int *devices, int *first_capture, int cap);
int sdr_multi_pcm_count(const sdr_multi *m, size_t nbytes_per_channel, size_t *n_pcm);
int sdr_multi_launch_count(sdr_multi *m, uint64_t *count, int reset);
/* iq [batch][iq_stride_bytes], pcm [batch][pcm_stride] in host memory (page-locked memory from
 * sdr_host_alloc is copied from / to directly; pageable memory is staged by the worker threads).
 * Returns when the whole batch's PCM is in `pcm`. */
int sdr_multi_process_host(sdr_multi *m, const uint8_t *iq, size_t iq_stride_bytes,
                           size_t nbytes_per_channel, int16_t *pcm, size_t pcm_stride);

/* ------------------------------------------------------------------------ */
/* Either side of the receiver.                   */
/* ------------------------------------------------------------------------ */
/* estimatePSD(freq, psd_est, samples, Fs)   include/fourier.h:29, src/fourier.cpp:44-126:
 * Bartlett estimate over 512-sample (include/dy4.h:27) Hann-windowed segments, in dB, averaged in
 * dB as the reference does.  samples: [rows][stride] floats in host memory, n per row (>= 512);
 * freq: 256 values (may be NULL); psd: [rows][256]. */
int sdr_psd(int device, const float *samples, size_t rows, size_t stride, size_t n, float Fs,
            float *freq, float *psd);
/* The same on intermediate `stage` (SDR_TAP_*) of the last process call of a pipeline, for every
 * capture, without leaving the device until the 256 values per capture come back (needs
 * sdr_pipeline_keep_taps).  Fs is the stage's own rate (IF rate or audio rate). */
int sdr_pipeline_psd(sdr_pipeline *p, int stage, float *freq, float *psd);

/* De-emphasis (time constant tau, 75e-6 in the Americas, 50e-6 elsewhere) on the receiver's PCM, in
 * place: the output stage the course spec left out (doc/3dy4-project-2022.pdf p.6).  One-pole
 * y[n] = (1-b) x[n] + b y[n-1], b = exp(-1/(Fs tau)); state per capture and audio channel carried
 * on the device.  pcm is [batch][pcm_stride] int16, n_frames frames of `channels` values. */
typedef struct sdr_deemph sdr_deemph;
int sdr_deemph_create(int device, int batch, int channels, float Fs, float tau, sdr_deemph **out);
int sdr_deemph_destroy(sdr_deemph *d);
int sdr_deemph_reset(sdr_deemph *d);
int sdr_deemph_process_device(sdr_deemph *d, int16_t *d_pcm, size_t pcm_stride, size_t n_frames, void *stream);
int sdr_deemph_process_host(sdr_deemph *d, int16_t *pcm, size_t pcm_stride, size_t n_frames);

/* 44-byte RIFF/WAVE header (16-bit PCM) for the receiver's output (48000 / 44100 Hz, 1 | 2 channels). */
int sdr_wav_header(uint8_t *out44, int sample_rate, int channels, uint64_t n_frames);

/* Polyphase channeliser in FRONT of the receiver: `n_wide` wideband captures, unsigned 8-bit I/Q at
 * n_channels x Fs, each become n_channels captures at Fs in device memory, in the format
 * sdr_pipeline_process_device consumes (row w * n_channels + c = the band centred c * Fs above the
 * wideband centre, wrapping: c > n_channels/2 lies below it).  Critically sampled analysis bank:
 * prototype = impulseResponseLPF at the wideband rate (n_channels * taps_per_branch taps, cut-off 0.8
 * of half the channel spacing), decimation n_channels, DFT across the branches; filter history is
 * carried on the device between calls. */
typedef struct sdr_channelizer sdr_channelizer;
typedef struct {
  int n_channels;      /* 2, 4, 8 or 16 */
  int taps_per_branch; /* 2..64 */
  int n_wide;          /* wideband captures per call */
  int device;
  float gain;          /* output scale before requantisation to 8 bits (0 = 1.0) */
} sdr_channelizer_config;
int sdr_channelizer_create(const sdr_channelizer_config *cfg, sdr_channelizer **out);
int sdr_channelizer_destroy(sdr_channelizer *c);
int sdr_channelizer_reset(sdr_channelizer *c);
int sdr_channelizer_prototype(const sdr_channelizer *c, float *h, size_t cap, size_t *n);
/* d_wide [n_wide][wide_stride] -> d_out [n_wide * n_channels][out_stride], nbytes_wide / n_channels bytes per row */
int sdr_channelizer_process_device(sdr_channelizer *c, const uint8_t *d_wide, size_t wide_stride,
                                   size_t nbytes_wide, uint8_t *d_out, size_t out_stride, void *stream);
int sdr_channelizer_process_host(sdr_channelizer *c, const uint8_t *wide, size_t wide_stride,
                                 size_t nbytes_wide, uint8_t *out, size_t out_stride);

/* ------------------------------------------------------------------------ */
/* RDS receiver chain (modes 0 and 2), attached to a batched pipeline.       */
/* The reference has no C++ RDS path: these entry points replace the block    */
/* loop of its Python model, model/fmRDS.py:222-276 (functions in             */
/* model/fmSupportLib.py), which works in double precision -- and so do the   */
/* device kernels (csrc/rds.cu).  Input is the pipeline's fm_demod.           */
/* Once created, the chain runs at the end of every sdr_pipeline_process_*    */
/* call of its pipeline (same stream); the pipeline's granule becomes one     */
/* RDS block (sdr_rds_info.block_bytes).  sdr_pipeline_reset resets it too.   */
/* Destroy the sdr_rds handle before its pipeline.  sdr_pipeline_copy_state   */
/* does not copy the RDS chain's state.                                       */
/* ------------------------------------------------------------------------ */
typedef struct sdr_rds sdr_rds;

typedef struct {
  int block_if;           /* IF samples per RDS block = the CDR window; 0 = the model's
                             (fmRDS.py:149-152: 9600 in mode 0, 1536000 in mode 2); a multiple of 9600 */
  int max_pending_blocks; /* blocks whose bits may wait for sdr_rds_read; 0 = default */
  int keep_nco;           /* keep the PLL's NCO outputs for sdr_rds_tap (SDR_RDS_TAP_PLL_*) */
  int cdr_carry;          /* 0: the CDR state is re-created for every block, as fmRDS.py:257-260 does;
                             1: it is carried from block to block (pair, start, prev_size:
                             fmSupportLib.py:104-106,178-189), so symbols are not dropped at block edges */
  int pll_form;           /* which form of the 114 kHz PLL kernel runs: SDR_RDS_PLL_AUTO picks by batch
                             size (a warp per capture up to 4096 captures, one lane per capture above);
                             both give the same result to 1e-9 (tests) */
  int precision;          /* SDR_RDS_F64 (default: the model's arithmetic, parity <= 1e-9) or SDR_RDS_F32_FIR:
                             the three FIR stages (channel / carrier band-pass, RRC) multiply and accumulate
                             in single precision (RRC output within 1e-5 of the model, SURVEY 8c's bar) */
} sdr_rds_config;
#define SDR_RDS_F64 0
#define SDR_RDS_F32_FIR 1
#define SDR_RDS_PLL_AUTO 0
#define SDR_RDS_PLL_LANE 1
#define SDR_RDS_PLL_WARP 2

typedef struct {
  int upsamp, decim;        /* fmRDS.py:57-58 / :69-70 */
  int samples_per_symbol;   /* fmRDS.py:60 / :72 */
  int block_if, block_out;  /* IF samples in / symbol-rate samples out per block */
  int block_bytes;          /* raw I/Q bytes per block */
  int max_pending_blocks, max_bits_per_block;
} sdr_rds_info_t;

/* Stages that sdr_rds_tap returns (last process call; names are fmRDS.py's variables). */
#define SDR_RDS_TAP_CHANNEL 0     /* fmRDS.py:223 rds_channel_filt */
#define SDR_RDS_TAP_CARRIER 1     /* fmRDS.py:233 rds_carrier_filt */
#define SDR_RDS_TAP_PLL_I 2       /* fmRDS.py:236 rds_PLL (n+1 values; needs keep_nco) */
#define SDR_RDS_TAP_PLL_Q 3       /* fmRDS.py:236 rds_PLL_Q */
#define SDR_RDS_TAP_MIXER_I 4     /* fmRDS.py:241 rds_mixer */
#define SDR_RDS_TAP_MIXER_Q 5     /* fmRDS.py:251 rds_mixer2 */
#define SDR_RDS_TAP_RESAMPLER_I 6 /* fmRDS.py:244 rds_resampler_out */
#define SDR_RDS_TAP_RESAMPLER_Q 7 /* fmRDS.py:252 rds_resampler_out2 */
#define SDR_RDS_TAP_RRC_I 8       /* fmRDS.py:248 rds_RRC_out */
#define SDR_RDS_TAP_RRC_Q 9       /* fmRDS.py:254 rds_RRC_out2 */

#define SDR_RDS_FILTER_CHANNEL 0   /* bandPass(151, if_fs, 54e3, 60e3)         fmRDS.py:122 */
#define SDR_RDS_FILTER_CARRIER 1   /* bandPass(151, if_fs, 113.5e3, 114.5e3)   fmRDS.py:123 */
#define SDR_RDS_FILTER_RESAMPLER 2 /* impResponse(101*U, if_fs*U, 3e3)         fmRDS.py:124 */
#define SDR_RDS_FILTER_RRC 3       /* impulseResponseRootRaisedCosine(2375*SPS, 101) fmRDS.py:125 */

/* The model's coefficient sets (host, double): fmSupportLib.py:358-385, :251-287. */
int sdr_rds_design(int which, int mode, double *h, size_t cap, size_t *n);

int sdr_rds_create(sdr_pipeline *p, const sdr_rds_config *cfg, sdr_rds **out);
int sdr_rds_destroy(sdr_rds *r);
int sdr_rds_info(const sdr_rds *r, sdr_rds_info_t *out);
/* Blocks processed since the last sdr_rds_discard. */
int sdr_rds_pending(sdr_rds *r, size_t *n_blocks);
/* Bit layer of capture `channel` for the pending blocks, in order:
 *   cdr_bits   CDR + Manchester decoding output      fmRDS.py:268 (fmSupportLib.py:103-222)
 *   diff_bits  differentially decoded bits           fmRDS.py:271 (fmSupportLib.py:241-249)
 *   bit_counts bits per block
 *   offsets    the frame synchroniser's result after each block: ' ', 'A', 'B', 'C',
 *              'c' (C') or 'D'                       fmRDS.py:272-276 (fmSupportLib.py:30-100)
 * Any output pointer may be NULL; *n_bits / *n_blocks always receive the totals. */
int sdr_rds_read(sdr_rds *r, int channel, uint8_t *cdr_bits, uint8_t *diff_bits, size_t bits_cap,
                 size_t *n_bits, int *bit_counts, char *offsets, size_t blocks_cap,
                 size_t *n_blocks);
/* Forget the pending blocks (the frame synchroniser keeps its carried bits). */
int sdr_rds_discard(sdr_rds *r);
/* Double-precision intermediate `stage` of capture `channel` from the last process call. */
int sdr_rds_tap(sdr_rds *r, int stage, int channel, double *dst, size_t cap, size_t *n);

#ifdef __cplusplus
}
#endif
#endif /* SDR_B200_H */
