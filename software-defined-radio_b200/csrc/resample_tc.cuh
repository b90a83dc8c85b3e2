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
// the slabs its blocks touch in time order through NST stages.  Three roles besides the 256
// read-back threads:
//  * producer (one thread): per slab step two TMA tensor copies bring the slab of both planes
//    (128 rows x 64 bytes each, 64-byte swizzle; rows past the batch arrive as zeros) and one bulk
//    copy brings the tap tiles of the <= 4 blocks active at that slab, which the host stored side
//    by side per slab position.  (The first version filled the stages with 16-byte cp.async from
//    all threads: 48 LDGSTS.128 per step at ~32 cycles each on the SM's load/store unit were the
//    kernel's whole run time, profiles/r2 -- and generic-proxy writes need a proxy fence before the
//    tensor core may read them; the asynchronous-proxy copies need neither.)
//    The same warp PLANS each step for the issuer (lane i decodes schedule entry i, two votes give the
//    runs of adjacent accumulator slots) and leaves an 80-byte record next to the stage: the issuer's
//    own instruction stream is the kernel's critical path, the producer mostly waits.
//  * issuer (one warp in lockstep, tcgen05 instructions under elect.sync so that ptxas emits them
//    without its ELECT / BRA.U.ANY serialisation loops): a small MMA costs a fixed 68 cycles whatever N <= 128 is
//    (tools/ubench_umma_smalln.cu), so MMAs are made as wide as the schedule allows: a block's
//    tile holds hh and hl side by side (32 accumulator columns: sum x*hh and sum x*hl, added in the
//    read-back), the active blocks' tiles are adjacent, and ONE MMA per K step and plane
//    (N = 32 x #blocks) serves them all: xh*[hh|hl] and xl*[hh|hl].  Every MMA accumulates.
//  * read-back (8 warps): one step after a block's last slab its accumulator (32 of 256
//    tensor-memory columns, a ring of 8 slots) is read, cleared for the next tenant, the two halves
//    added, scaled, converted to int16 and stored.
#pragma once

#include <cuda.h>   // CUtensorMap (the encode entry point is fetched at run time: no libcuda link)

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
constexpr size_t rt_smem(int nst) { return (size_t)nst * RT_STAGE + 1024; }   // + alignment of the swizzled tiles

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
  uint32_t any_last[RT_MAX_SP];       // != 0: some block has its last slab at q
  uint32_t nact[RT_MAX_SP];           // entries at q; their tap tiles are stored side by side:
  uint32_t bq_off[RT_MAX_SP];         // [chunk 4][32 * nact rows][16 B] at tile offset bq_off[q] of `tiles`
};

struct ResampleTcArgs {
  int pl_off;                         // element of a plane row that holds sample 0 (history before it); the planes
                                      // themselves come as tensor maps
  const uint8_t *tiles;               // per slab position q: [4 chunks][32 rows per active block: hh, hl of its 16 outputs][8 halfs]
  int16_t *pcm;                       // [B][pcm_stride]
  size_t pcm_stride;
  float *audio_filt;                  // optional [B][tap_stride]
  size_t tap_stride;
  float out_scale;                    // 2^-S
  int batch;
  int n_periods;                      // periods in this call
  int ctas_per_tile;                  // CTAs sharing one 128-capture tile (they split its blocks)
};

#ifdef SDR_RT_TRACE
// Debug build only: cycles one CTA's roles spend in each wait (printed by CTA 7).
#define RT_T0() const long long rt_t0__ = clock64()
#define RT_ACC(v) (v) += clock64() - rt_t0__
#else
#define RT_T0() do { } while (0)
#define RT_ACC(v) do { } while (0)
#endif

// The period tables travel as a kernel parameter (constant bank): the per-step schedule lookups are
// constant-cache reads.  NST pipeline stages, MINB resident CTAs per SM.
template <int NST, int MINB>
static __global__ void __launch_bounds__(RT_BLOCK, MINB)
k_audio_resample_tc(const ResampleTcArgs g, const __grid_constant__ RtTables tab,
                    const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_l) {
  extern __shared__ __align__(128) uint8_t rt_smem_raw[];
  const uint32_t smem_u32 = (tc_smem_u32(rt_smem_raw) + 1023u) & ~1023u;   // swizzled tiles want their natural alignment
  __shared__ __align__(8) uint64_t full[NST], empty[NST], acc_full[RT_SLOTS], acc_empty[RT_SLOTS];
  __shared__ uint32_t tmem_slot;
  // One step's work for the issuer, planned by the producer thread (which runs NST steps ahead and
  // spends most of its time waiting): the issuer's own instruction stream was the kernel's critical path
  // -- ~1950 cycles per slab step in one thread, of which ~850 decoded the schedule (debug build
  // SDR_RT_TRACE) -- so it now only reads this record, waits, and issues.
  struct __align__(16) Cmd {
    uint32_t n_run;            // MMA runs (1 or 2): consecutive accumulator slots served by one MMA per K step and plane
    uint32_t b_lbo;            // B descriptor's leading-dimension offset (32 * nact rows * 16 B) >> 4
    uint32_t d[2], b_off[2], idesc[2];   // per run: accumulator address offset, B start (address units), instruction descriptor
    uint32_t wait_slot[RT_NACT], wait_par[RT_NACT];   // per schedule entry: its slot; parity | 2 when the block starts here and must wait for the slot's previous tenant
    uint32_t done_slot[RT_NACT];                      // per schedule entry: slot | 8 when the block ends here
  };
  __shared__ Cmd cmds[NST];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef SDR_RT_TRACE
  long long tr_prod_empty = 0, tr_iss_full = 0, tr_iss_acc = 0, tr_rb_full = 0;
  const long long tr_begin = clock64();
#endif

  if (tid == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(&full[i], 1);    // the producer's arrive.expect_tx; the copies complete the bytes
      mbar_init(&empty[i], 1);   // tcgen05.commit of the MMAs that read the stage
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
  pdl_trigger();
  pdl_wait();   // the planes come from the front end's kernel; nothing above read them
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
      int lb, b, p;       // block index relative to gb0, block within its period, period
      bool first, last, valid;
    };
    auto active = [&](int p, int q, Ent (&e)[RT_NACT]) {   // entry i keeps position i (static indexing)
      const uint4 w4 = *reinterpret_cast<const uint4 *>(tab.sched[q]);
      const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
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
        e[i].first = ((w >> 8) & 0xff) == 0;
        e[i].last = (w >> 17) & 1;
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

    if (warp == RT_WORKERS / 32 + 1) {
      // ---- producer: asynchronous-proxy copies into the stage the MMAs of NST steps ago have read ----
      {
        // the whole warp plans (lane i < 4 decodes schedule entry i; votes give the runs), lane 0 copies
        const uint32_t idesc16 = (1u << 4) | ((uint32_t)(RT_ROWS >> 4) << 24);   // f16 x f16 -> f32, M = 128
        int p = p_first, q = q_first;
        for (int st = 0; st < n_steps; ++st, advance(p, q)) {
          const int stage = st % NST;
          if (st >= NST && lane == 0) { RT_T0(); mbar_wait(&empty[stage], ((st / NST) - 1) & 1); RT_ACC(tr_prod_empty); }
          __syncwarp();
          // ---- plan the issuer's step (the previous tenant of this record was consumed NST steps ago) ----
          {
            Cmd &c = cmds[stage];
            const uint32_t w = lane < RT_NACT ? tab.sched[q][lane] : 0u;
            const int bb = (int)(w & 0xff);
            const int pp = p + (int)((w >> 16) & 1);
            const int gb = pp * NBLK + bb;
            const bool valid = (w >> 18) && gb >= gb0 && gb < gb1;
            const int lb = gb - gb0, slot = lb & (RT_SLOTS - 1);
            const bool first = ((w >> 8) & 0xff) == 0, last = (w >> 17) & 1;
            const uint32_t vmask = __ballot_sync(0xffffffffu, valid) & ((1u << RT_NACT) - 1);
            // a run of consecutive accumulator slots starts at the first valid entry and where the ring wraps
            const bool start = valid && (lane == 0 || !((vmask >> (lane - 1)) & 1) || slot == 0);
            const uint32_t smask = __ballot_sync(0xffffffffu, start) & ((1u << RT_NACT) - 1);
            if (lane < RT_NACT) {
              c.wait_slot[lane] = (uint32_t)slot;
              c.wait_par[lane] = (valid && first && lb >= RT_SLOTS) ? (uint32_t)((((lb / RT_SLOTS) - 1) & 1) | 2u) : 0u;   // bit 1: wait
              c.done_slot[lane] = (valid && last) ? (uint32_t)slot | 8u : 0u;                                                 // bit 3: commit
              if (start) {
                const uint32_t above = ~((2u << lane) - 1);                       // entries after this one
                const uint32_t stop = ((smask | ~vmask) & above) | (1u << RT_NACT);   // next start, first gap, or the end
                const int cnt = (__ffs(stop) - 1) - lane;
                const int r = __popc(smask & ((1u << lane) - 1));
                c.d[r] = (uint32_t)(slot * RT_NC);
                c.b_off[r] = (uint32_t)(lane * RT_NC);          // entry i's tile: 32 rows of 16 bytes = 32 address units
                c.idesc[r] = idesc16 | ((uint32_t)((RT_NC * cnt) >> 3) << 17);
              }
            }
            if (lane == 0) {
              c.n_run = (uint32_t)__popc(smask);
              c.b_lbo = tab.nact[q] * RT_NC;
            }
          }
          __syncwarp();
          if (lane == 0) {
            const uint32_t dst = smem_u32 + (uint32_t)stage * RT_STAGE, bar = tc_smem_u32(&full[stage]);
            const uint32_t nb = tab.nact[q] * RT_TILE_BYTES;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)RT_X_BYTES + nb) : "memory");
            const int x0 = g.pl_off + (T0 + st) * RT_SLAB;
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                         "l"(&map_h), "r"(x0), "r"(c0), "r"(bar)
                         : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst + RT_X_BYTES / 2),
                         "l"(&map_l), "r"(x0), "r"(c0), "r"(bar)
                         : "memory");
            if (nb)
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + RT_X_BYTES),
                           "l"(g.tiles + (size_t)tab.bq_off[q] * RT_TILE_BYTES), "r"(nb), "r"(bar)
                           : "memory");
          }
        }
      }
      __syncwarp();
    } else if (warp == RT_WORKERS / 32) {
      // ---- issuer: executes the records the producer planned.  The whole warp walks the steps in
      // lockstep and ONE ELECTED lane issues: under a plain `if (lane == 0)` ptxas wraps every
      // tcgen05 instruction in an ELECT / BRA.U.ANY loop over "possibly different" operand values ----
      {
        // Matrix descriptors, K-major.  A (a plane's slab: 128 rows of 64 bytes, 64-byte swizzle as the
        // tensor copy wrote it): 8-row groups 512 B apart, layout type 4; a K step of 16 halfs is 32 bytes
        // further into the swizzle atom.  B (tap tiles, no swizzle, 16-byte core-matrix rows): 8-row groups
        // 128 B apart, chunks of K (32 * nact rows) * 16 B apart.  Version 1 in both.
        constexpr uint32_t A_HI = (512u >> 4) | (1u << 14) | (4u << 29), B_HI = (128u >> 4) | (1u << 14);
        for (int st = 0; st < n_steps; ++st) {
          const int stage = st % NST;
          { RT_T0(); mbar_wait(&full[stage], (st / NST) & 1); RT_ACC(tr_iss_full); }
          asm volatile("tcgen05.fence::after_thread_sync;");
          const Cmd &c = cmds[stage];   // read in place (shared memory): no dynamically indexed local copy
          const uint32_t xs_lo = ((smem_u32 + (uint32_t)stage * RT_STAGE) >> 4) & 0x3fff;
          // a block that starts here takes over an accumulator slot: its previous tenant must have
          // been read back (and cleared)
          const uint32_t n_run = c.n_run, b_lbo = c.b_lbo;
          const uint4 wp = *reinterpret_cast<const uint4 *>(c.wait_par), ws4 = *reinterpret_cast<const uint4 *>(c.wait_slot);
          const uint4 dn = *reinterpret_cast<const uint4 *>(c.done_slot);
          const uint32_t wpar[RT_NACT] = {wp.x, wp.y, wp.z, wp.w}, wslot[RT_NACT] = {ws4.x, ws4.y, ws4.z, ws4.w};
          const uint32_t done[RT_NACT] = {dn.x, dn.y, dn.z, dn.w};
#pragma unroll
          for (int i = 0; i < RT_NACT; ++i)
            if (wpar[i] & 2u) {
              { RT_T0(); mbar_wait(&acc_empty[wslot[i]], wpar[i] & 1u); RT_ACC(tr_iss_acc); }
              asm volatile("tcgen05.fence::after_thread_sync;");
            }
#ifdef SDR_RT_TRACE
          const long long tr_m = clock64();
#endif
          const uint32_t bs_lo = (((smem_u32 + (uint32_t)stage * RT_STAGE + RT_X_BYTES) >> 4) & 0x3fff) | (b_lbo << 16);
          if (tc_elect_one()) {
          for (uint32_t r = 0; r < n_run; ++r) {
            const uint32_t d = tmem + c.d[r], b_lo = bs_lo + c.b_off[r], idesc = c.idesc[r];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
              for (int plane = 0; plane < 2; ++plane) {      // xh, xl
                const uint32_t a_lo = xs_lo + plane * ((RT_X_BYTES / 2) >> 4) + ks * (32 >> 4);
                const uint32_t bb_lo = b_lo + ks * 2 * b_lbo;
                asm volatile(
                    "{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %4};\nmov.b64 db, {%2, %5};\n"
                    "setp.ne.b32 p, %5, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}\n" ::"r"(d),
                    "r"(a_lo), "r"(bb_lo), "r"(idesc), "r"(A_HI), "r"(B_HI));
              }
            }
          }
#ifdef SDR_RT_TRACE
          const long long tr_c = clock64();
          tr_rb_full += tr_c - tr_m;   // (issuer: MMA issue)
#endif
          tc_commit(&empty[stage]);            // the stage may be refilled once these MMAs have read it
#pragma unroll
          for (int i = 0; i < RT_NACT; ++i)
            if (done[i] & 8u) tc_commit(&acc_full[done[i] & 7u]);
#ifdef SDR_RT_TRACE
          tr_prod_empty += clock64() - tr_c;   // (issuer: commits)
#endif
          }
          __syncwarp();
        }
      }
      __syncwarp();
    } else {
      // ---- read-back: finished accumulators -> PCM ----
      const int half = warp >> 2;                                  // which 8 of a block's 16 outputs
      const uint32_t tlane = (uint32_t)((warp & 3) * 32) << 16;    // this warp's quarter of the accumulator lanes
      const int ocap = c0 + (warp & 3) * 32 + lane;                // accumulator row = capture
      int p = p_first, q = q_first;
      for (int st = 0; st < n_steps; ++st, advance(p, q)) {
        if (tab.any_last[q] == 0) continue;
        Ent e[RT_NACT];
        active(p, q, e);
#pragma unroll
        for (int k = 0; k < RT_NACT; ++k) {
          if (!e[k].valid || !e[k].last) continue;
          const int lb = e[k].lb, slot = lb % RT_SLOTS;
          { RT_T0(); mbar_wait(&acc_full[slot], (lb / RT_SLOTS) & 1); RT_ACC(tr_rb_full); }
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
            const int j0 = e[k].b * RT_NB + half * 8;             // first of this thread's outputs in the period
            const int nv = min(8, P_out - j0);                    // valid outputs (the last block of a period is partial)
            const long long o0 = (long long)e[k].p * P_out + j0;
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
      }
    }
  }
#ifdef SDR_RT_TRACE
  if (blockIdx.x == 7 && lane == 0 && (warp == 0 || warp >= RT_WORKERS / 32))
    printf("rt trace cta 7 warp %d: total %lld  producer waits empty %lld  issuer waits full %lld, acc_empty %lld  read-back waits acc_full %lld\n",
           warp, clock64() - tr_begin, tr_prod_empty, tr_iss_full, tr_iss_acc, tr_rb_full);
#endif
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(RT_TMEM_COLS));
}

}  // namespace sdr
