// resample_tc.cuh -- tensor-core polyphase resampler (SDR_VARIANT_FAST, mono, modes 2/3).
//
// Replaces convolveBlockResampleFIR (filter.cpp:191-223) + the PCM conversion
// (threadMonoOnly.cpp:185-190) on the fast path.  Why: the CUDA-core quad resampler moves one
// shared-memory word per multiply-add pair and sits at its data-pipe roof (profiles/r1l: 0.245 ms
// of a 0.68 ms step for 6 % of the step's arithmetic).
//
// The resampler is a PERIODIC banded matrix: with g = gcd(U, D), every P_in = D/g input samples
// produce P_out = U/g outputs, and output j of a period uses the TA inputs ending at i0 = j*D/U
// with the taps of phase (j*D) % U.  Cut a period's outputs into blocks of 16 and the time axis
// into slabs of 32 samples; block b meets slab q through a fixed 16 x 32 tap tile H[b][q] (zero
// where an output's window does not reach).  With 128 captures as the rows of the A operand,
//     Y[128 captures x 16 outputs] += X[128 x 32 (slab q)] * H[b][q]^T
// is one tcgen05.mma (kind::f16, fp32 accumulators in tensor memory) per 16 samples of K.
//
// Precision.  fm_demod and the taps are split into two halves, x = xh + xl and h' = hh + hl
// (h' = h * (1+U) * 2^S: the reference's output gain, filter.cpp:213, and a power-of-two scale that
// puts the taps in the fp16 range are folded in), each half an fp16 with an 11-bit significand:
//     y * 2^S = sum xh*hh + sum xh*hl + sum xl*hh + sum xl*hl
// Every product is exact in the fp32 accumulator; the result differs from the reference's
// sequential float sum by ~1e-6 relative (tests: audio >= 100 dB, PCM +-1 LSB).  The front end
// writes the two planes itself (rf_tc.cuh), so no conversion pass runs here.
//
// Schedule ("K outer").  A CTA owns 128 captures x a range of consecutive output blocks and walks
// the slabs its blocks touch in time order.  Per slab step the workers copy the slab of both
// planes (2 x 8 KB, canonical no-swizzle K-major core-matrix layout: chunk kc of row r at
// kc*2048 + r*16) and the tap tiles of the <= 4 blocks active at that slab (concatenated along N)
// into one of NST stages with 16-byte asynchronous copies.  A small MMA costs a fixed 68 cycles
// whatever N <= 128 is (tools/ubench_umma_smalln.cu), so MMAs are made as wide as the schedule
// allows: a block's tap tile holds hh and hl side by side (32 accumulator columns: sum x*hh and
// sum x*hl, added in the read-back) and the tiles of all blocks active at a slab sit next to each
// other, so ONE MMA per K step and plane (N = 32 x #blocks) serves them all: xh*[hh|hl] and
// xl*[hh|hl] (which also brings the xl*hl term).  Every MMA accumulates; the workers clear an
// accumulator (32 of 256 tensor-memory columns, a ring of 8 slots) right after reading it back,
// one step after the block's last slab: add the two halves, scale, int16 conversion, PCM store.
#pragma once

#include "rf_tc.cuh"

namespace sdr {

constexpr int RT_ROWS = 128;          // captures per CTA (= accumulator lanes)
constexpr int RT_NB = 16;             // outputs per block (accumulator columns per block)
constexpr int RT_SLAB = 32;           // input samples per slab (two K steps of 16)
constexpr int RT_NACT = 4;            // blocks active at one slab, at most
constexpr int RT_SLOTS = 8;           // accumulator slots in tensor memory
constexpr int RT_NST = 4;             // pipeline stages (x 2 resident CTAs per SM)
constexpr int RT_WORKERS = 256;
constexpr int RT_BLOCK = RT_WORKERS + 64;   // + the MMA issue warp + the proxy-fence warp
constexpr int RT_X_BYTES = 2 * (RT_SLAB / 8) * RT_ROWS * 16;           // both planes of one slab: 16 KB
constexpr int RT_NC = 2 * RT_NB;      // accumulator columns per block: sum x*hh | sum x*hl
constexpr int RT_BROWS = RT_NC * RT_NACT;                              // 128 tap-tile rows per stage
constexpr int RT_B_BYTES = (RT_SLAB / 8) * RT_BROWS * 16;              // [hh | hl] tiles of the active blocks: 8 KB
constexpr int RT_TMEM_COLS = RT_SLOTS * RT_NC;                         // 256
constexpr int RT_STAGE = RT_X_BYTES + RT_B_BYTES;
constexpr int RT_TILE_BYTES = (RT_SLAB / 8) * RT_NC * 16;              // one (block, slab) tile in global memory: 2 KB
constexpr int RT_MAX_SP = 128;        // slabs per period, at most (mode 3: 100)
constexpr int RT_MAX_BLK = 64;        // blocks per period, at most (mode 3: 28)
constexpr size_t rt_smem(int nst) { return (size_t)nst * RT_STAGE + 1024; }

// Host-built description of one period (uploaded once per pipeline).
struct RtTables {
  int P_in, P_out, SP, NBLK;          // samples in, outputs out, slabs, blocks per period
  int qmin;                           // first (negative) slab position a period's blocks reach back to
  int qs[RT_MAX_BLK], qe[RT_MAX_BLK]; // slab range of block b, relative to its period's first sample
  int tile0[RT_MAX_BLK];              // index of block b's first tile in the tile array
  // Blocks that meet slab position q of a period (in ascending block order): those of the same
  // period (dp = 0) and those of the next one that reach back to it (dp = 1).  One word per
  // entry: b | slab ordinal << 8 | dp << 16 | last slab of the block << 17 | valid << 18.
  alignas(16) uint32_t sched[RT_MAX_SP][RT_NACT];
  alignas(16) uint32_t tile[RT_MAX_SP][RT_NACT];  // tile index (tile0[b] + slab ordinal) of the same entries
  uint32_t any_last[RT_MAX_SP];       // != 0: some block has its last slab at q
};

struct ResampleTcArgs {
  const uint16_t *xh, *xl;            // [B][pl_stride] fp16 planes, sample 0 at pl_off (history before it)
  size_t pl_stride;
  int pl_off;
  const uint8_t *tiles;               // [sum_b nslab_b][4 chunks][32 rows: hh of 16 outputs, hl of 16 outputs][8 halfs]
  int16_t *pcm;                       // [B][pcm_stride]
  size_t pcm_stride;
  float *audio_filt;                  // optional [B][tap_stride]
  size_t tap_stride;
  float out_scale;                    // 2^-S
  int batch;
  int n_periods;                      // periods in this call
  int ctas_per_tile;                  // CTAs sharing one 128-capture tile (they split its blocks)
};

// The period tables travel as a kernel parameter (constant bank): the per-step schedule lookups are
// then constant-cache reads instead of shared-memory loads queued behind the asynchronous copies
// (profiles/r2h: short-scoreboard stalls on exactly those loads).
// NST pipeline stages, MINB resident CTAs per SM.  FENCER: a third role (one thread of warp 9)
// waits for a stage's copies, executes the generic->async proxy fence and only then releases the
// stage to the MMA issuer, which otherwise pays for that fence between its MMAs.  Measured on the
// bench workload (tools/exp_rt_variants.sh, round 2): 4 stages x 2 CTAs 0.206 ms, with the fence
// warp 0.194; 8 stages x 1 CTA 0.243 / 0.188.  tools/ubench_umma_smalln.cu: one of these MMAs
// (M 128, K 16, operands in shared memory) costs 68 cycles whatever N <= 128 is.
template <int NST, int MINB, bool FENCER>
static __global__ void __launch_bounds__(RT_BLOCK, MINB)
k_audio_resample_tc(const ResampleTcArgs g, const __grid_constant__ RtTables tab) {
  extern __shared__ __align__(128) uint8_t rt_smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(rt_smem_raw) + 127) & ~(uintptr_t)127);
  __shared__ __align__(8) uint64_t full[NST], ready[NST], empty[NST], acc_full[RT_SLOTS], acc_empty[RT_SLOTS];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool issuer = warp == RT_WORKERS / 32;

  if (tid == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(&full[i], RT_WORKERS);
      mbar_init(&ready[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < RT_SLOTS; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], RT_WORKERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_slot)), "n"(RT_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  // every MMA accumulates: start from cleared accumulators (each worker warp clears its lane
  // quarter of half of the columns)
  if (warp < RT_WORKERS / 32) {
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * (RT_TMEM_COLS / 2);
#pragma unroll
    for (int c = 0; c < RT_TMEM_COLS / 2; c += 8)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(base + c), "r"(0u));
    asm volatile("tcgen05.wait::st.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  // (registers: every asm with a memory clobber would otherwise make the compiler re-read them from shared memory)
  const int NBLK = tab.NBLK, SP = tab.SP, P_out = tab.P_out;

  // ---- this CTA's range of global blocks gb = p * NBLK + b of its capture tile ----
  const int rt = blockIdx.x / g.ctas_per_tile, part = blockIdx.x % g.ctas_per_tile;
  const int total_blocks = g.n_periods * NBLK;
  const int gb0 = (int)((long long)total_blocks * part / g.ctas_per_tile);
  const int gb1 = (int)((long long)total_blocks * (part + 1) / g.ctas_per_tile);
  const int c0 = rt * RT_ROWS;
  if (gb0 < gb1) {
    // absolute slab positions T = p * SP + q walked by this CTA
    const int T0 = (gb0 / NBLK) * SP + tab.qs[gb0 % NBLK];
    const int T1 = ((gb1 - 1) / NBLK) * SP + tab.qe[(gb1 - 1) % NBLK];
    const int n_steps = T1 - T0 + 1;

    // Blocks active at slab position q of period p (T = p * SP + q; p = -1 for the history slabs in
    // front of period 0), restricted to this CTA's range: read from the host-built schedule.
    struct Ent {
      int lb, b, js, p;   // block index relative to gb0, block within its period, slab ordinal, period
      uint32_t tile;      // index of the (block, slab) tap tile
      bool last, valid;
    };
    auto active = [&](int p, int q, Ent (&e)[RT_NACT]) {   // entry i keeps position i (static indexing)
      const uint4 w4 = *reinterpret_cast<const uint4 *>(tab.sched[q]);
      const uint4 t4 = *reinterpret_cast<const uint4 *>(tab.tile[q]);
      const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w}, ts[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
      for (int i = 0; i < RT_NACT; ++i) {
        const uint32_t w = ws[i];
        const int bb = (int)(w & 0xff);
        const int pp = p + (int)((w >> 16) & 1);
        const int gb = pp * NBLK + bb;
        e[i].valid = (w >> 18) && gb >= gb0 && gb < gb1;
        e[i].lb = gb - gb0;
        e[i].p = pp;
        e[i].b = bb;
        e[i].js = (int)((w >> 8) & 0xff);
        e[i].last = (w >> 17) & 1;
        e[i].tile = ts[i];
      }
    };
    // (p, q) of step 0; every role then advances its own pair by one slab per step
    const int p_first = T0 >= 0 ? T0 / SP : -1;
    const int q_first = T0 - p_first * SP;
    auto advance = [&](int &p, int &q) {
      if (++q == SP) {
        q = 0;
        ++p;
      }
    };

    if (issuer) {
      if (lane == 0) {
        const uint32_t idesc16 = (1u << 4) | ((uint32_t)(RT_ROWS >> 4) << 24);   // f16 x f16 -> f32, M = 128
        const uint32_t smem_u32 = tc_smem_u32(smem);
        int p = p_first, q = q_first;
#ifdef SDR_RT_TRACE
        long long tr_wait = 0, tr_fence = 0, tr_issue = 0, tr_t0 = clock64();
#endif
        for (int st = 0; st < n_steps; ++st, advance(p, q)) {
          const int stage = st % NST;
          Ent e[RT_NACT];
          active(p, q, e);
#ifdef SDR_RT_TRACE
          const long long tr_a = clock64();
#endif
          if (FENCER) {
            mbar_wait(&ready[stage], (st / NST) & 1);
          } else {
            mbar_wait(&full[stage], (st / NST) & 1);
#ifdef SDR_RT_TRACE
            tr_wait += clock64() - tr_a;
#endif
            // the workers' asynchronous copies (generic proxy) have landed: order them before the tensor
            // core's reads (async proxy)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          }
          asm volatile("tcgen05.fence::after_thread_sync;");
#ifdef SDR_RT_TRACE
          const long long tr_c = clock64();
#endif
          // Matrix descriptors (no swizzle, K-major): the high word is the same for every operand of a
          // kind (stride between 8-row groups 128 B, version 1); the low word is the 16-byte address
          // unit plus the K-direction stride, so operands are told apart by small additions.
          constexpr uint32_t HI = (128u >> 4) | (1u << 14);
          constexpr uint32_t A_LBO = ((uint32_t)(RT_ROWS * 16) >> 4) << 16, B_LBO = ((uint32_t)(RT_BROWS * 16) >> 4) << 16;
          constexpr uint32_t A_KS = (2 * RT_ROWS * 16) >> 4, A_PLANE = ((RT_SLAB / 8) * RT_ROWS * 16) >> 4;
          constexpr uint32_t B_KS = (2 * RT_BROWS * 16) >> 4;
          const uint32_t xs_lo = ((smem_u32 + (uint32_t)stage * RT_STAGE) >> 4) | A_LBO;
          const uint32_t bs_lo = ((smem_u32 + (uint32_t)stage * RT_STAGE + RT_X_BYTES) >> 4) | B_LBO;
          // a block that starts here takes over an accumulator slot: its previous tenant must have
          // been read back (and cleared)
#pragma unroll
          for (int i = 0; i < RT_NACT; ++i)
            if (e[i].valid && e[i].js == 0 && e[i].lb >= RT_SLOTS) {
              mbar_wait(&acc_empty[e[i].lb % RT_SLOTS], ((e[i].lb / RT_SLOTS) - 1) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;");
            }
          // one MMA per K step and plane for every run of consecutive accumulator slots (entry i's
          // tile sits at rows 32 i of the stage, its accumulator in slot lb % 8: a run ends where the
          // ring of slots wraps)
#pragma unroll
          for (int i = 0; i < RT_NACT; ++i) {
            if (!e[i].valid) continue;
            const int slot0 = e[i].lb % RT_SLOTS;
            if (i > 0 && e[i - 1].valid && slot0 != 0) continue;   // part of the run that started earlier
            int cnt = 1;
#pragma unroll
            for (int k = 1; k < RT_NACT; ++k)
              if (i + k < RT_NACT && cnt == k && e[(i + k) & (RT_NACT - 1)].valid && slot0 + k < RT_SLOTS) cnt = k + 1;
            const uint32_t idesc = idesc16 | ((uint32_t)((RT_NC * cnt) >> 3) << 17);
            const uint32_t d = tmem + slot0 * RT_NC;
            const uint32_t b_lo = bs_lo + (uint32_t)(i * RT_NC);   // 32 rows of 16 bytes = 32 address units
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
              for (int plane = 0; plane < 2; ++plane) {      // xh, xl
                const uint32_t a_lo = xs_lo + plane * A_PLANE + ks * A_KS;
                const uint32_t bb_lo = b_lo + ks * B_KS;
                asm volatile(
                    "{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %4};\nmov.b64 db, {%2, %4};\n"
                    "setp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}\n" ::"r"(d),
                    "r"(a_lo), "r"(bb_lo), "r"(idesc), "r"(HI));
              }
            }
          }
          tc_commit(&empty[stage]);            // the stage may be refilled once these MMAs have read it
#pragma unroll
          for (int k = 0; k < RT_NACT; ++k)
            if (e[k].valid && e[k].last) tc_commit(&acc_full[e[k].lb % RT_SLOTS]);
#ifdef SDR_RT_TRACE
          tr_fence += tr_c - tr_a;
          tr_issue += clock64() - tr_c;
#endif
        }
#ifdef SDR_RT_TRACE
        if (blockIdx.x == 5)
          printf("issuer: steps %d total %lld wait_full %lld wait+fence %lld issue %lld\n", n_steps, clock64() - tr_t0, tr_wait,
                 tr_fence, tr_issue);
#endif
      }
      __syncwarp();
    } else if (warp == RT_WORKERS / 32 + 1) {
      if (FENCER && lane == 0) {
        for (int st = 0; st < n_steps; ++st) {
          const int stage = st % NST;
          mbar_wait(&full[stage], (st / NST) & 1);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(&ready[stage])) : "memory");
        }
      }
      __syncwarp();
    } else {
      // ---- workers: fill stages with asynchronous copies, read finished accumulators back ----
      // Slab copies: a row contributes 64 contiguous bytes per plane, so four lanes take one row (one
      // 16-byte chunk each) and a warp instruction covers eight rows = eight 128-byte lines (a lane per
      // row would touch 32 lines per instruction and saturate the L1 tag stage: profiles/r2e).
      // Thread t: plane t / 128; warp w of the plane: rows 32 w + 8 i + lane / 4 (i = 0..3), chunk lane % 4.
      const int plane = tid >> 7, kc = lane & 3;
      const int row0 = ((tid >> 5) & 3) * 32 + (lane >> 2);
      const uint16_t *plane_base = (plane ? g.xl : g.xh) + g.pl_off + 8 * kc;
      size_t row_off[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)   // rows past the batch shadow the last capture
        row_off[i] = (size_t)min(c0 + row0 + 8 * i, g.batch - 1) * g.pl_stride;
#ifdef SDR_RT_TRACE
      long long trw_empty = 0, trw_rb = 0, trw_t0 = clock64();
#endif
      int pl = p_first, ql = q_first;   // slab position of the next step to be loaded
      auto load = [&](int st) {
        if (st < n_steps) {
          const int T = T0 + st, stage = st % NST;
          Ent e[RT_NACT];
          active(pl, ql, e);   // (shared-memory reads first: they would queue behind the copies below)
          advance(pl, ql);
#ifdef SDR_RT_TRACE
          const long long tr_e = clock64();
#endif
          if (st >= NST) mbar_wait(&empty[stage], ((st / NST) - 1) & 1);
#ifdef SDR_RT_TRACE
          trw_empty += clock64() - tr_e;
#endif
          uint8_t *xs = smem + (size_t)stage * RT_STAGE;
          const uint32_t xdst = tc_smem_u32(xs) + plane * ((RT_SLAB / 8) * RT_ROWS * 16) + kc * RT_ROWS * 16 + row0 * 16;
          const uint16_t *src = plane_base + (long long)T * RT_SLAB;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            // ask L2 for the whole 256-byte neighbourhood, so that DRAM sees one long burst per row
            // instead of four short ones over the next slabs
            asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;" ::"r"(xdst + i * 8 * 16),
                         "l"(src + row_off[i])
                         : "memory");
          const uint32_t bdst = tc_smem_u32(xs) + RT_X_BYTES;
          // tile chunk w = kc * 32 + row  ->  stage offset (kc * BROWS + 32 a + row) * 16;
          // a thread copies chunk w = tid % 128 of entries a = tid / 128 and a + 2
          const int w = tid & (RT_TILE_BYTES / 16 - 1), wk = w / RT_NC, wr = w % RT_NC;
#pragma unroll
          for (int a = 0; a < RT_NACT; ++a) {
            if ((a & 1) != (tid >> 7) || !e[a].valid) continue;
            const uint8_t *tsrc = g.tiles + (size_t)e[a].tile * RT_TILE_BYTES + (size_t)w * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(bdst + (wk * RT_BROWS + a * RT_NC + wr) * 16),
                         "l"(tsrc)
                         : "memory");
          }
        }
        // this thread's arrival on the stage's barrier fires when its copies above have landed
        // (no thread ever waits for its own copies)
        if (st < n_steps)
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc_smem_u32(&full[st % NST])) : "memory");
      };
      for (int st = 0; st < NST - 1; ++st) load(st);
      const int half = warp >> 2;                                  // which 8 of a block's 16 outputs
      const uint32_t tlane = (uint32_t)((warp & 3) * 32) << 16;    // this warp's quarter of the accumulator lanes
      const int arow = (warp & 3) * 32 + lane;                     // accumulator row = capture within the tile
      const int ocap = c0 + arow;
      int pr = p_first, qr = q_first;   // slab position of the next step to be read back
      auto read_back = [&]() {   // blocks whose last slab was that step
        const bool some = tab.any_last[qr] != 0;
        Ent e[RT_NACT];
        if (some) active(pr, qr, e);
        advance(pr, qr);
        if (!some) return;
#pragma unroll
        for (int k = 0; k < RT_NACT; ++k) {
          if (!e[k].valid || !e[k].last) continue;
          const int b = e[k].b, p = e[k].p;
          const int lb = e[k].lb, slot = lb % RT_SLOTS;
          mbar_wait(&acc_full[slot], (lb / RT_SLOTS) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;");
          uint32_t v[8], u[8];
          const uint32_t tcol = tmem + tlane + slot * RT_NC + half * 8;
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                       : "r"(tcol));
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                       : "r"(tcol + RT_NB));
          asm volatile("tcgen05.wait::ld.sync.aligned;");
          // clear both halves for the slot's next tenant (every MMA accumulates)
          asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tcol), "r"(0u));
          asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tcol + RT_NB), "r"(0u));
          asm volatile("tcgen05.wait::st.sync.aligned;");
          asm volatile("tcgen05.fence::before_thread_sync;");
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(&acc_empty[slot])) : "memory");
          if (ocap < g.batch) {
            const int j0 = b * RT_NB + half * 8;                  // first of this thread's outputs in the period
            const int nv = min(8, P_out - j0);                // valid outputs (the last block of a period is partial)
            const long long o0 = (long long)p * P_out + j0;
            int16_t *dst = g.pcm + (size_t)ocap * g.pcm_stride + o0;
            float y[8];
            int16_t s[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              y[i] = __fmul_rn(__fadd_rn(__uint_as_float(v[i]), __uint_as_float(u[i])), g.out_scale);
              s[i] = pcm16(y[i]);
            }
            if (nv == 8 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                reinterpret_cast<uint32_t *>(dst)[i] = (uint32_t)(uint16_t)s[2 * i] | ((uint32_t)(uint16_t)s[2 * i + 1] << 16);
            } else if (nv == 8) {   // starts on an odd sample: one half-word, three words, one half-word
              dst[0] = s[0];
#pragma unroll
              for (int i = 0; i < 3; ++i)
                reinterpret_cast<uint32_t *>(dst + 1)[i] = (uint32_t)(uint16_t)s[2 * i + 1] | ((uint32_t)(uint16_t)s[2 * i + 2] << 16);
              dst[7] = s[7];
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (i < nv) dst[i] = s[i];
            }
            if (g.audio_filt) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (i < nv) g.audio_filt[(size_t)ocap * g.tap_stride + o0 + i] = y[i];
            }
          }
        }
      };
      // Per step: refill the stage that step st - 1's MMAs have read (with the slab of step
      // st + NST - 1) and read back the accumulators they finished.
      for (int st = 0; st < n_steps; ++st) {
        load(st + NST - 1);
#ifdef SDR_RT_TRACE
        const long long tr_r = clock64();
#endif
        if (st > 0) read_back();
#ifdef SDR_RT_TRACE
        trw_rb += clock64() - tr_r;
#endif
      }
      read_back();
#ifdef SDR_RT_TRACE
      if (blockIdx.x == 5 && tid == 0)
        printf("worker0: total %lld wait_empty %lld read_back %lld\n", clock64() - trw_t0, trw_empty, trw_rb);
#endif
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(RT_TMEM_COLS));
}

}  // namespace sdr
