// rf_tc.cuh -- tensor-core RF front end (SDR_VARIANT_FAST, mono; rf_decim D = 10, 5 or 3).
//
// Why: the bit-exact CUDA-core front end is FP32-issue bound (profiles/r1a: the FMA pipe is the
// busiest unit, DRAM below 10 %): 30.2 multiply-adds per input sample against 2 bytes, and the
// exact form needs two instructions per multiply-add.  north_star admits a Toeplitz GEMM in
// exactly that situation.
//
// How (polyphase Hankel GEMM on tcgen05, kind::i8), for a T-tap filter decimating by D:
//   y[j] = sum_n h[n] x[D j - n] = sum_p sum_q h[D q + p - c] * xp[p][j - q],   xp[p][i] = x[D i - p + c]
// (c = TcCfg::SHIFT, a free re-indexing that aligns the transposer's windows)
// 1. A CUDA-core pass transposes the raw interleaved bytes into 2 D byte streams (D phases x
//    {I,Q}) in shared memory -- no conversion, the bytes stay unsigned 8-bit (PRMT only).
// 2. For each stream, row m of the MMA's A operand is the K bytes starting 16 bytes after row
//    m-1: a matrix descriptor with leading-byte-offset 16 and stride-byte-offset 128 turns the
//    stream into that overlapping-row (Hankel) matrix in place (tools/umma_hankel_test.cu), so
//    row m sees xp[p][j0+16m-(BACK-FRONT) ..] and produces the 16 outputs j0+16m+delta.
// 3. The B operand of phase p holds that phase's taps, scaled to 31-bit fixed point and split
//    into four signed base-256 digits: column 4*delta+d has digit_d at k = delta+(BACK-FRONT)-q.
//    (Three digits are enough for 130 dB on I/Q in steady state, but not while the filter fills
//    at the start of a capture, where the outputs are ~1e-6 and fmDemod divides by them.)
//    D*K/32 MMAs per component accumulate sum_t digit_d(h[t]) * u8[...] exactly in int32; one
//    more MMA with a constant A operand starts the accumulators at -128 * sum_t digit_d(h[t]),
//    so they end as the digit sums over the CENTRED samples (x - 128), each below 2^22.
// 4. The epilogue recombines the four digit sums (tc_combine): exact in integers up to two
//    31-bit halves, then two int->float conversions and one FMA -- at most one ulp from the
//    exactly rounded fixed-point FIR output (tap quantisation 2^-34, far below the reference's
//    own float rounding), and exactly zero when that output is zero.  It differs from the
//    reference's sequential float sum by ~1e-7 relative (tests: >= 100 dB SNR, PCM +-1 LSB;
//    against an integer model of the kernel: <= 1 ulp).
//    fmDemod follows in the same kernel; the one-sample state crosses rows via shared memory.
//
// Schedule: persistent CTAs, two per SM, 8 worker warps + 1 issue warp.  Work items (capture,
// segment of tiles) come from a device counter.  Per tile the workers transpose the staged bytes
// into the streams, arrive on an mbarrier and go on to the epilogue of the PREVIOUS tile
// (accumulators are double buffered; a TMEM lane quarter is readable by warps w and w+4, which
// split a row's 16 outputs); lane 0 of the issue warp waits for the 256 arrivals, starts the TMA
// bulk copy of the next tile into the (now free) staging buffer and then issues the MMAs.
// History of the design with the measured time of each step: DESIGN.md section 4, profiles/,
// tools/tc_trace.py.
#pragma once

#include "kernels.cuh"

namespace sdr {

constexpr int TC_ROWS = 128;                 // A rows per tile (= TMEM lanes)
constexpr int TC_THREADS = 2 * TC_ROWS;       // worker threads (transpose + epilogue)
constexpr int TC_BLOCK = TC_THREADS + 32;    // + one warp that only issues MMAs and bulk copies
constexpr int TC_TILE_OUT = TC_ROWS * 16;    // 2048 outputs per tile
constexpr int TC_TMAX = 151;                 // longest supported filter
constexpr int TC_ND = 4;                     // signed base-256 digits per tap (31-bit fixed point)
constexpr int TC_N = 16 * TC_ND;             // 64 accumulator columns per component
constexpr int TC_CORR_BYTES = TC_N * 32;      // one 64 x 32 B tile that removes the +128 offset (below)
constexpr int TC_MAX_SEGS = 15;              // segments per capture (work items per capture)
constexpr int TC_FRONT = 16;                 // spare stream entries in front of row 0's window

template <int D>
struct TcCfg {
  static constexpr int Q = (TC_TMAX + D - 1) / D;              // taps per phase: 16 / 31 / 51
  static constexpr int K = (16 + Q - 1 + 31) / 32 * 32;        // bytes per A row: 32 / 64 / 96
  static constexpr int KSTEPS = K / 32;
  // Two free choices make every transposer group's input window start on a 16-byte boundary
  // (one LDS.128 less per group): the phase split is shifted by SHIFT taps (stream p holds
  // x[D i - p + SHIFT], so tap t sits at n = t + SHIFT = D q + p), and stream entry s' holds
  // xp[p][j0 - BACK + s'] with any BACK in [FRONT + Q - 1, FRONT + K - 16].
  static constexpr int SHIFT = D == 10 ? 1 : D == 5 ? 2 : 0;
  static constexpr int BACK = D == 10 ? 32 : TC_FRONT + Q - 1;
  static_assert(BACK >= TC_FRONT + Q - 1 && BACK <= TC_FRONT + K - 16 && TC_TMAX - 1 + SHIFT <= D * Q - 1, "stream geometry");
  static constexpr int STREAM = TC_FRONT + 16 * (TC_ROWS - 1) + K;
  static constexpr int NSTREAM = 2 * D;
  static constexpr int GE = (D % 2 == 0) ? 4 : 8;              // stream entries built per transposer group
  static constexpr int NGRP = STREAM / GE;
  static constexpr int WB = 2 * D * GE;                        // input bytes consumed per group
  static constexpr int CONST0 = 2 * D * BACK + 2 * (D - 1) - 2 * SHIFT;    // 2*c_lo of group 0 = 2*D*j0 - CONST0
  static constexpr int OFF = (16 - CONST0 % 16) % 16;          // bytes from the aligned window start
  static constexpr int WIN = (WB + OFF + 15) / 16 * 16;        // aligned window read per group
  static constexpr int BASE = CONST0 + OFF;                    // window of group 0 starts at 2*D*j0 - BASE
  static constexpr int RAW_BYTES = WB * (NGRP - 1) + WIN;      // staged bytes per tile
  static constexpr int RAW = (RAW_BYTES + 127) / 128 * 128;
  static constexpr int BP = TC_N * K;                          // bytes of one phase's B tile
  static constexpr int HIST = (BASE + 1) / 2;                  // raw history pairs the first tile reaches back
  static constexpr int NTAPQ = D * Q;                          // fixed-point tap table length
  static constexpr size_t SMEM = (size_t)RAW + (size_t)NSTREAM * STREAM + (size_t)D * BP + TC_CORR_BYTES + 128;
  static_assert(STREAM % GE == 0 && WB % 16 == 0 && BASE % 16 == 0, "tile geometry");
};

struct RfTcArgs {
  RfArgs a;
  const int8_t *bmat;   // [D phases][64 x K] in canonical no-swizzle K-major core-matrix order, then
                        // the 64 x 32 offset tile
  const int32_t *hq;    // fixed-point taps, D*Q entries (zero padded)
  float scale;          // 2^-(S+7)
  // work item = one segment (consecutive tiles [seg_begin[s], seg_begin[s+1])) of one capture;
  // items are numbered segment-major, so the short segments at the end of a capture are the last
  // items handed out and the kernel's tail (CTAs waiting for the slowest) stays short
  int seg_begin[TC_MAX_SEGS + 1];
  int segs, batch;      // items = segs * batch
  int *next_item;       // device counter, zeroed before the launch: further items are gridDim.x + atomicAdd
};

__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(tc_smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
// True in exactly one lane of a converged warp (elect.sync).
__device__ __forceinline__ bool tc_elect_one() {
  uint32_t p;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
      tc_smem_u32(bar)) : "memory");
}

// 32 accumulator columns (8 outputs x 4 digits) of this warp's 32 TMEM lanes.
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
      "%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

// Recombine the four base-256 digit sums (already centred: the offset tile subtracted
// 128 * sum(digit) inside the tensor core, so each |d_k| < 2^22) into the fixed-point FIR output
// v = d0 + 2^8 d1 + 2^16 d2 + 2^24 d3 and scale it.  lo = d0 + 2^8 d1 and hi = d2 + 2^8 d3 are
// exact 31-bit integers; each is rounded to float once and fma(hi, 2^16 sc, lo sc) rounds a
// third time: at most one ulp from the exactly rounded v * sc, and exactly 0 for v = 0
// (profiles/r1h: the exactly rounded form was a quarter of the kernel's instructions).
struct TcScale {
  float sc;     // 2^-(S+7)
  float sc16;   // 2^16 * sc
};
__device__ __forceinline__ float tc_combine(const uint32_t *d, const TcScale &k) {
  const int lo = (int32_t)d[0] + 256 * (int32_t)d[1];
  const int hi = (int32_t)d[2] + 256 * (int32_t)d[3];
  return __fmaf_rn(__int2float_rn(hi), k.sc16, __fmul_rn(__int2float_rn(lo), k.sc));
}
// fmDemod on the fast path: same formula, approximate reciprocal (the fast variant is held to
// 100 dB / +-1 LSB against the reference, not to bit equality; I/Q already differ by ~1e-7).
__device__ __forceinline__ float tc_demod(float i, float q, float pi, float pq) {
  const float den = __fmaf_rn(i, i, __fmul_rn(q, q));
  const float num = __fmaf_rn(i, __fsub_rn(q, pq), -__fmul_rn(q, __fsub_rn(i, pi)));
  float r;   // I, Q are multiples of 2^-(S+7): den is 0 or far above the subnormal range
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
  return den == 0.0f ? 0.0f : __fmul_rn(num, r);
}

// One transposer group: GE consecutive entries of all 2*D streams from D*GE input pairs that sit
// OFF bytes into the aligned window `w`.  Entry ds of phase p is pair (D-1) + D*ds - p.
template <int D>
__device__ __forceinline__ void tc_transpose_group(const uint32_t *w, uint8_t *streams, int grp) {
  using C = TcCfg<D>;
#pragma unroll
  for (int p = 0; p < D; ++p) {
#pragma unroll
    for (int comp = 0; comp < 2; ++comp) {
#pragma unroll
      for (int h = 0; h < C::GE / 4; ++h) {
        const int P0 = 2 * ((D - 1) + D * (4 * h) - p) + comp + C::OFF;
        const int P1 = P0 + 2 * D, P2 = P0 + 4 * D, P3 = P0 + 6 * D;
        const uint32_t ab = __byte_perm(w[P0 >> 2], w[P1 >> 2], (P0 & 3) | ((4 + (P1 & 3)) << 4));
        const uint32_t cd = __byte_perm(w[P2 >> 2], w[P3 >> 2], (P2 & 3) | ((4 + (P3 & 3)) << 4));
        *reinterpret_cast<uint32_t *>(streams + (2 * p + comp) * C::STREAM + C::GE * grp + 4 * h) =
            __byte_perm(ab, cd, 0x5410);
      }
    }
  }
}

#ifdef SDR_TC_TRACE
// Debug build only (tools/tc_trace.py): clock64 stamps of one CTA, [warp 0..8][tile 0..15][stamp 0..11].
__device__ long long g_tc_trace[9 * 16 * 12];
__device__ long long g_tc_begin[1024], g_tc_end[1024];
#define TC_STAMP(it, k)                                                                         \
  do {                                                                                          \
    if (tc_trace_on && blockIdx.x == 7 && lane == 0 && (it) < 16)                           \
      g_tc_trace[(warp * 16 + (it)) * 12 + (k)] = clock64();                                    \
  } while (0)
#else
#define TC_STAMP(it, k) do { } while (0)
#endif

template <int D>
static __global__ void __launch_bounds__(TC_BLOCK, 2)
k_rf_demod_tc(const RfTcArgs g) {
  using C = TcCfg<D>;
  const RfArgs &a = g.a;
  extern __shared__ __align__(128) uint8_t tc_smem[];
  uint8_t *raw = tc_smem;                                   // [C::RAW] staged input bytes
  uint8_t *streams = tc_smem + C::RAW;                      // [2D][C::STREAM]
  int8_t *bs = reinterpret_cast<int8_t *>(tc_smem + C::RAW + C::NSTREAM * C::STREAM);
  int8_t *corr_b = bs + D * C::BP;                          // [64 x 32] offset tile (B operand)
  uint8_t *corr_a = reinterpret_cast<uint8_t *>(corr_b) + TC_CORR_BYTES;   // one core matrix of 128s (A operand)
  // I,Q of delta 7 / 15 of every row, [tile % 3][half][row]: a set is rewritten three tiles later,
  // i.e. after two more worker barriers, so readers of tile t and t+1 are always done with it
  __shared__ float last_i[3][2][TC_ROWS], last_q[3][2][TC_ROWS];
  __shared__ float carry_iq[2][2];                      // [round parity][I,Q]
  __shared__ int red[2][2][TC_ND][TC_THREADS / 32];     // [round parity][I,Q][digit][warp]
  __shared__ __align__(8) uint64_t mma_done[2], raw_full, streams_ready, item_full, item_empty;
  __shared__ int item_slot[2];   // item of the CTA's k-th round at [k & 1], -1 = no more work
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  bool tc_trace_on = true;   // used by the SDR_TC_TRACE debug build only
  TC_STAMP(15, 10);
#ifdef SDR_TC_TRACE
  long long tc_begin_ns;
  {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    tc_begin_ns = (long long)ns;
  }
#endif
  const int rowi = (warp & 3) * 32 + lane;  // TMEM lane = A row served by this thread
  const int half = warp >> 2;               // which 8 of the row's 16 outputs
  const int n_tiles = (a.n_if + TC_TILE_OUT - 1) / TC_TILE_OUT;
  const int n_items = g.segs * g.batch;     // work item = (segment, capture)
  const bool issuer = warp == TC_THREADS / 32;   // warp 8: lane 0 issues MMAs and bulk copies
  auto workers_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(TC_THREADS) : "memory"); };

  // ---- once per (persistent) CTA: tap digits, barriers, tensor memory ----
  for (int i = tid; i < (D * C::BP + TC_CORR_BYTES) / 16; i += TC_BLOCK)
    reinterpret_cast<uint4 *>(bs)[i] = __ldg(reinterpret_cast<const uint4 *>(g.bmat) + i);
  if (tid < 32) reinterpret_cast<uint32_t *>(corr_a)[tid] = 0x80808080u;
  if (tid == 0) {
    mbar_init(&mma_done[0], 1);
    mbar_init(&mma_done[1], 1);
    mbar_init(&raw_full, 1);
    mbar_init(&streams_ready, TC_THREADS);
    mbar_init(&item_full, 1);
    mbar_init(&item_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(tc_smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");  // bs written by the generic proxy
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  // everything above read constants only; the input history, the carried I/Q and the work counter
  // come from the previous call's k_carry
  pdl_trigger();
  pdl_wait();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) |
                         ((uint32_t)(TC_ROWS >> 4) << 24);
  const TcScale ks{g.scale, g.scale * 65536.0f};

  // Geometry of a work item.  A tile that lies entirely inside the capture is staged by ONE bulk
  // asynchronous copy (TMA, completes on raw_full); tiles that touch the history in front of the
  // capture or its end are assembled chunk by chunk by the workers.
  struct Item {
    int b, tile_begin, tile_end;
    const uint8_t *row, *hrow;
    bool row_aligned;
  };
  auto get_item = [&](int item) {
    Item w;
    const int sg = item / g.batch;
    w.b = item - sg * g.batch;
    w.tile_begin = g.seg_begin[sg];
    w.tile_end = g.seg_begin[sg + 1];
    w.row = a.iq + (size_t)w.b * a.iq_stride;
    w.hrow = a.hist + (size_t)w.b * 2 * a.rf_hist_len;
    w.row_aligned = ((reinterpret_cast<uintptr_t>(w.row) & 15) == 0);
    return w;
  };
  auto tile_is_bulk = [&](const Item &w, int tile) -> bool {
    const long long wbase = 2ll * D * tile * TC_TILE_OUT - C::BASE;
    return w.row_aligned && wbase >= 0 && wbase + C::RAW_BYTES <= 2 * a.n_rf;
  };
  auto issue_bulk = [&](const Item &w, int tile) {   // one thread
    const long long wbase = 2ll * D * tile * TC_TILE_OUT - C::BASE;
    constexpr uint32_t BYTES = C::RAW_BYTES;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(&raw_full)), "r"(BYTES)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            tc_smem_u32(raw)),
        "l"(w.row + wbase), "r"(BYTES), "r"(tc_smem_u32(&raw_full))
        : "memory");
  };
  auto stage_chunked = [&](const Item &w, int tile) {  // all workers
    const long long wbase = 2ll * D * tile * TC_TILE_OUT - C::BASE;
    for (int q = tid; q < C::RAW_BYTES / 16; q += TC_THREADS) {
      const long long pos = wbase + 16ll * q;
      if (w.row_aligned && pos >= 0 && pos + 16 <= 2 * a.n_rf) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc_smem_u32(raw + 16 * q)), "l"(w.row + pos)
                     : "memory");
      } else if (pos >= 2 * a.n_rf) {
        *reinterpret_cast<uint4 *>(raw + 16 * q) = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
      } else {
        // edge chunk: 16 independent byte loads (one latency), centred zero beyond either end
        uint32_t v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const long long p = pos + k;
          const long long h = 2ll * a.rf_hist_len + p;
          const uint8_t *src = p < 0 ? (h >= 0 ? w.hrow + h : nullptr) : (p < 2 * a.n_rf ? w.row + p : nullptr);
          v[k] = src ? (uint32_t)__ldg(src) : 128u;
        }
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = v[4 * k] | (v[4 * k + 1] << 8) | (v[4 * k + 2] << 16) | (v[4 * k + 3] << 24);
        *reinterpret_cast<uint4 *>(raw + 16 * q) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // Work items.  Round r of a CTA processes item(r): item(0) = blockIdx.x, item(r + 1) is taken
  // from the device counter by worker 0 at the START of round r (dynamic balance: CTA run times
  // differ by +-12 %, tools/tc_trace.py) and published in item_slot[(r + 1) & 1]: phase r of
  // item_full.  The issue warp acknowledges on item_empty so that phases never alias.
  // Tiles of consecutive items are pipelined like tiles of one item: the first tile of
  // item(r + 1) is staged and transposed before the last epilogue of item(r).
  if (issuer) {
    // ---- issue warp.  Per tile: wait until all workers have written its streams (which also
    // means `raw` is free), start the bulk copy of the NEXT tile -- it goes first because issuing
    // the MMAs blocks for most of their run time (tools/tc_trace.py) -- then the MMAs. ----
    if (lane == 0) {
      uint32_t gt = 0;  // tiles issued so far by this CTA
      Item w = get_item(blockIdx.x);
      for (uint32_t round = 0;; ++round) {
        tc_trace_on = round == 2;
        int next_item = -1;
        Item nw = w;
        for (int tile = w.tile_begin; tile < w.tile_end; ++tile, ++gt) {
          const uint32_t buf = gt & 1;
          mbar_wait(&streams_ready, gt & 1);
          asm volatile("tcgen05.fence::after_thread_sync;");
          TC_STAMP(tile - w.tile_begin, 10);
          if (tile + 1 < w.tile_end) {
            if (tile_is_bulk(w, tile + 1)) issue_bulk(w, tile + 1);
          } else {
            mbar_wait(&item_full, round & 1);
            next_item = item_slot[(round + 1) & 1];
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(&item_empty)) : "memory");
            if (next_item >= 0) {
              nw = get_item(next_item);
              if (tile_is_bulk(nw, nw.tile_begin)) issue_bulk(nw, nw.tile_begin);
            }
          }
          const uint32_t s0 = tc_smem_u32(streams) + TC_FRONT, b0 = tc_smem_u32(bs);
          const uint32_t d0 = tmem + buf * 2 * TC_N;
#pragma unroll
          for (int comp = 0; comp < 2; ++comp) {
            {
              // offset tile first (overwrites the accumulators): every A row is the same 32 bytes
              // of 128 (both strides 0), column (delta, d) of B sums to -sum_t digit_d(h[t]), so the
              // accumulators start at -128 * sum_t digit_d(h[t]) and end as sums over (x - 128)
              const uint64_t da = tc_desc(tc_smem_u32(corr_a), 0, 0);
              const uint64_t db = tc_desc(tc_smem_u32(corr_b), (TC_N / 8) * 128, 128);
              asm volatile(
                  "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                  "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d0 + comp * TC_N),
                  "l"(da), "l"(db), "r"(idesc), "r"(0u));
            }
#pragma unroll
            for (int p = 0; p < D; ++p) {
#pragma unroll
              for (int ksx = 0; ksx < C::KSTEPS; ++ksx) {
                const uint64_t da = tc_desc(s0 + (2 * p + comp) * C::STREAM + 32 * ksx, 16, 128);
                const uint64_t db = tc_desc(b0 + p * C::BP + ksx * 2 * (TC_N / 8) * 128, (TC_N / 8) * 128, 128);
                const uint32_t acc = 1u;
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d0 + comp * TC_N),
                    "l"(da), "l"(db), "r"(idesc), "r"(acc));
              }
            }
          }
          tc_commit(&mma_done[buf]);
          TC_STAMP(tile - w.tile_begin, 11);
        }
        if (next_item < 0) break;
        w = nw;
      }
    }
    __syncwarp();
  } else {
    // ---- workers: transpose tile t, then the epilogue of tile t-1 while the MMAs of t run ----
    // Epilogue of tile `tile` of capture b (segment starting at tile_begin; global tile number t:
    // accumulator set t & 1, its (t >> 1)-th use; rp = parity of the item's round).
    auto epilogue = [&](int b, int tile_begin, int tile, uint32_t t, uint32_t rp) {
      const long long j0 = (long long)tile * TC_TILE_OUT;
      const uint32_t buf = t & 1;
      TC_STAMP(tile - tile_begin + 1, 4);
      mbar_wait(&mma_done[buf], (t >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;");
      TC_STAMP(tile - tile_begin + 1, 5);
      float fi[8], fq[8];
      const uint32_t trow = tmem + buf * 2 * TC_N + 32 * half + ((uint32_t)((warp & 3) * 32) << 16);
      {
        uint32_t v[32];
        tc_ld32(trow, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
        for (int k = 0; k < 8; ++k) fi[k] = tc_combine(&v[4 * k], ks);
        TC_STAMP(tile - tile_begin + 1, 6);
        tc_ld32(trow + TC_N, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
        for (int k = 0; k < 8; ++k) fq[k] = tc_combine(&v[4 * k], ks);
      }
      const int set = t % 3, pset = (set + 2) % 3;
      last_i[set][half][rowi] = fi[7];
      last_q[set][half][rowi] = fq[7];
      asm volatile("tcgen05.fence::before_thread_sync;");
      TC_STAMP(tile - tile_begin + 1, 7);
      workers_sync();
      TC_STAMP(tile - tile_begin + 1, 8);
      float pi, pq;
      if (half == 1) {            // delta 8 follows delta 7 of the same row
        pi = last_i[set][0][rowi];
        pq = last_q[set][0][rowi];
      } else if (rowi > 0) {      // delta 0 follows delta 15 of the previous row
        pi = last_i[set][1][rowi - 1];
        pq = last_q[set][1][rowi - 1];
      } else if (tile > tile_begin) {  // ... or the last output of the segment's previous tile
        pi = last_i[pset][1][TC_ROWS - 1];
        pq = last_q[pset][1][TC_ROWS - 1];
      } else if (tile != 0) {     // first output of a later segment: recomputed in integers
        uint32_t di[TC_ND], dq[TC_ND];
#pragma unroll
        for (int d = 0; d < TC_ND; ++d) {
          int si = 0, sq = 0;
          for (int k = 0; k < TC_THREADS / 32; ++k) {
            si += red[rp][0][d][k];
            sq += red[rp][1][d][k];
          }
          di[d] = (uint32_t)si;
          dq[d] = (uint32_t)sq;
        }
        pi = tc_combine(di, ks);   // the same function of the same digit sums as the tile that owns it
        pq = tc_combine(dq, ks);
      } else {                    // first output of the call: the carried state
        pi = carry_iq[rp][0];
        pq = carry_iq[rp][1];
      }
      const long long jrow = j0 + 16 * rowi + 8 * half;
      float dm[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        dm[k] = tc_demod(fi[k], fq[k], pi, pq);
        pi = fi[k];
        pq = fq[k];
      }
      if (a.xh && jrow + 8 <= a.n_if) {
        // split planes for the tensor-core resampler: 16 bytes per plane and thread (pl_off and
        // pl_stride are multiples of 8, jrow is a multiple of 8)
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint16_t h0, h1, l0, l1;
          float f0, f1;
          asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h0) : "f"(dm[2 * k]));
          asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h1) : "f"(dm[2 * k + 1]));
          asm("cvt.f32.f16 %0, %1;" : "=f"(f0) : "h"(h0));
          asm("cvt.f32.f16 %0, %1;" : "=f"(f1) : "h"(h1));
          asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(l0) : "f"(__fsub_rn(dm[2 * k], f0)));
          asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(l1) : "f"(__fsub_rn(dm[2 * k + 1], f1)));
          hw[k] = (uint32_t)h0 | ((uint32_t)h1 << 16);
          lw[k] = (uint32_t)l0 | ((uint32_t)l1 << 16);
        }
        const size_t at = (size_t)b * a.pl_stride + a.pl_off + jrow;
        *reinterpret_cast<uint4 *>(a.xh + at) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4 *>(a.xl + at) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
      float *drow = a.demod + (size_t)b * a.demod_stride + a.demod_off;
      if (!a.write_f32) {
        // the planes are the only consumer's input
      } else if (jrow + 8 <= a.n_if && ((a.demod_stride | a.demod_off) & 7) == 0) {
        // one 256-bit store per thread: half the store wavefronts of two float4 (the L1 data pipe
        // is the busiest unit of this kernel, profiles/r1k)
        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(drow + jrow), "f"(dm[0]), "f"(dm[1]),
                     "f"(dm[2]), "f"(dm[3]), "f"(dm[4]), "f"(dm[5]), "f"(dm[6]), "f"(dm[7])
                     : "memory");
      } else if (jrow + 8 <= a.n_if && ((a.demod_stride | a.demod_off) & 3) == 0) {
        *reinterpret_cast<float4 *>(drow + jrow) = make_float4(dm[0], dm[1], dm[2], dm[3]);
        *reinterpret_cast<float4 *>(drow + jrow + 4) = make_float4(dm[4], dm[5], dm[6], dm[7]);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (jrow + k < a.n_if) drow[jrow + k] = dm[k];
      }
      if (a.i_filt) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (jrow + k < a.n_if) {
            a.i_filt[(size_t)b * a.tap_stride + jrow + k] = fi[k];
            a.q_filt[(size_t)b * a.tap_stride + jrow + k] = fq[k];
          }
      }
      {
        const long long last = (long long)a.n_if - 1 - jrow;
        if (last >= 0 && last < 8) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k == last) {
              a.prev_out[2 * b] = fi[k];
              a.prev_out[2 * b + 1] = fq[k];
            }
        }
      }
      TC_STAMP(tile - tile_begin + 1, 9);
    };

    uint32_t gt = 0;         // tiles transposed so far by this CTA (same count as the issue warp's)
    uint32_t raw_phase = 0;  // completions of raw_full consumed so far
    uint32_t round = 0;
    Item w = get_item(blockIdx.x);
    int tile = w.tile_begin;
    int p_b = 0, p_begin = 0, p_tile = -1;   // previous tile (p_tile < 0: none yet)
    uint32_t p_rp = 0;
    if (!tile_is_bulk(w, tile)) stage_chunked(w, tile);
    else if (tid == 0) issue_bulk(w, tile);
    for (;;) {
      const int it = tile - w.tile_begin;
      if (it == 0) {   // start of a round
        tc_trace_on = true;
        TC_STAMP(15, round < 10 ? round : 9);
        tc_trace_on = round == 2;
        if (tid == 0) {
          if (round > 0) mbar_wait(&item_empty, (round - 1) & 1);   // the issue warp has read item(round)
          int nxt = (int)gridDim.x + atomicAdd(g.next_item, 1);
          if (nxt >= n_items) nxt = -1;
          item_slot[(round + 1) & 1] = nxt;
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(&item_full)) : "memory");
        }
        if (w.tile_begin == 0 && tid < 2) carry_iq[round & 1][tid] = a.prev_in[2 * w.b + tid];
      }
      const bool bulk = tile_is_bulk(w, tile);
      TC_STAMP(it, 0);
      if (bulk) {                                        // this tile's bytes have landed
        mbar_wait(&raw_full, raw_phase & 1);
        ++raw_phase;
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      TC_STAMP(it, 1);
      if (gt > 0) mbar_wait(&mma_done[(gt - 1) & 1], ((gt - 1) >> 1) & 1);   // MMAs of the previous tile have read `streams`
      TC_STAMP(it, 2);
      if (!bulk) workers_sync();                         // chunked staging: everyone's copies are visible
      // ---- 1. transpose raw -> 2 D phase streams ----
      for (int grp = tid; grp < C::NGRP; grp += TC_THREADS) {
        uint32_t wd[C::WIN / 4];
        const uint4 *src = reinterpret_cast<const uint4 *>(raw + C::WB * grp);
#pragma unroll
        for (int k = 0; k < C::WIN / 16; ++k) {
          const uint4 v = src[k];
          wd[4 * k] = v.x; wd[4 * k + 1] = v.y; wd[4 * k + 2] = v.z; wd[4 * k + 3] = v.w;
        }
        tc_transpose_group<D>(wd, streams, grp);
      }
      TC_STAMP(it, 3);
      // Every worker publishes its stream entries to the async proxy and arrives; only the
      // issue warp waits for all 256 arrivals -- the workers go on to the previous epilogue.
      asm volatile("fence.proxy.async.shared::cta;");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(&streams_ready)) : "memory");
      // ---- 2. which tile comes next (possibly the first tile of the next item) ----
      Item nw = w;
      int ntile = tile + 1;
      bool have_next = true;
      if (ntile == w.tile_end) {
        mbar_wait(&item_full, round & 1);
        const int nxt = item_slot[(round + 1) & 1];
        have_next = nxt >= 0;
        if (have_next) {
          nw = get_item(nxt);
          ntile = nw.tile_begin;
        }
      }
      const bool next_chunked = have_next && !tile_is_bulk(nw, ntile);
      const bool need_red = it == 0 && tile != 0;
      // rare: the first tile of a later segment reads other threads' stream entries below; a
      // chunked next tile overwrites `raw`
      if (need_red || next_chunked) workers_sync();
      if (need_red) {   // predecessor of the segment's first output: its centred digit sums, in integers
        int si[TC_ND] = {}, sq[TC_ND] = {};
        for (int n = tid; n < C::NTAPQ; n += TC_THREADS) {  // tap Dq+p meets stream entry BACK-1-q
          const int q = n / D, p = n - q * D;
          const int xi = (int)streams[(2 * p) * C::STREAM + C::BACK - 1 - q] - 128;
          const int xq = (int)streams[(2 * p + 1) * C::STREAM + C::BACK - 1 - q] - 128;
          int v = g.hq[n];
#pragma unroll
          for (int d = 0; d < TC_ND; ++d) {
            const int digit = ((v + 128) & 255) - 128;   // balanced base-256 digit, as in the B tiles
            v = (v - digit) >> 8;
            si[d] += digit * xi;
            sq[d] += digit * xq;
          }
        }
#pragma unroll
        for (int d = 0; d < TC_ND; ++d) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            si[d] += __shfl_xor_sync(0xffffffffu, si[d], o);
            sq[d] += __shfl_xor_sync(0xffffffffu, sq[d], o);
          }
          if (lane == 0) {   // read after the worker barrier of this tile's epilogue
            red[round & 1][0][d][warp] = si[d];
            red[round & 1][1][d][warp] = sq[d];
          }
        }
      }
      if (next_chunked) stage_chunked(nw, ntile);   // a bulk next tile is issued by the issue warp
      // ---- 3. epilogue of the previous tile while this tile's MMAs run ----
      if (p_tile >= 0) epilogue(p_b, p_begin, p_tile, gt - 1, p_rp);
      else workers_sync();   // every iteration has one barrier after the item_full read above
      p_b = w.b;
      p_begin = w.tile_begin;
      p_tile = tile;
      p_rp = round & 1;
      ++gt;
      if (!have_next) break;
      if (ntile == nw.tile_begin && tile + 1 == w.tile_end) ++round;
      w = nw;
      tile = ntile;
    }
    epilogue(p_b, p_begin, p_tile, gt - 1, p_rp);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
  tc_trace_on = true;
  TC_STAMP(15, 11);
#ifdef SDR_TC_TRACE
  if (lane == 0 && warp == 0) {   // wall-clock end of every CTA (ns) and of the traced CTA's start
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    g_tc_end[blockIdx.x] = (long long)ns;
    g_tc_begin[blockIdx.x] = tc_begin_ns;
  }
#endif
}

}  // namespace sdr
