// rf_tc.cuh -- tensor-core RF front end (SDR_VARIANT_FAST, mono, rf_decim = 10).
//
// Why: the bit-exact CUDA-core front end is FP32-issue bound (profiles/: the FMA pipe is the
// busiest unit, DRAM below 10 %): 30.2 multiply-adds per input sample against 2 bytes, and
// the exact form needs two instructions per multiply-add.  north_star admits a Toeplitz
// GEMM in exactly that situation.
//
// How (polyphase Hankel GEMM on tcgen05, kind::i8):
//   y[j] = sum_n h[n] x[10j - n] = sum_p sum_q h[10q+p] * xp[p][j-q],   xp[p][i] = x[10i - p]
// 1. A CUDA-core pass transposes the raw interleaved bytes into 20 byte streams (10 phases x
//    {I,Q}) in shared memory -- no conversion, the bytes stay unsigned 8-bit.
// 2. For each stream, row m of the MMA's A operand is the 32 bytes starting 16 bytes after
//    row m-1: a matrix descriptor with leading-byte-offset 16 and stride-byte-offset 128
//    turns the stream into that overlapping-row (Hankel) matrix in place, so row m sees
//    xp[p][j0+16m-15 .. j0+16m+16] and produces the 16 outputs j0+16m+delta.
// 3. The B operand of phase p holds the taps h[10q+p], scaled to 23-bit fixed point and
//    split into three signed base-256 digits: column 3*delta+d has digit_d at k = delta+15-q.
//    Ten MMAs (one per phase) accumulate sum_t digit_d(h[t]) * u8[...] exactly in int32.
// 4. The epilogue recombines the digits in int64, removes the 128 offset of the unsigned
//    samples and rounds ONCE to float: the exactly rounded fixed-point FIR output (tap
//    quantisation 2^-26, i.e. below the reference's own float rounding).  It differs from the reference's sequential float sum only by the
//    reference's own accumulated rounding (~1e-7 relative; >= 100 dB SNR, PCM +-1 LSB).
//    fmDemod follows in the same kernel with the reference's float operations; the
//    one-sample state travels between rows by warp shuffle.
// Per 128-row tile: 2048 outputs = 20480 input pairs, 20 MMAs of 128 x 48 x 32.
#pragma once

#include "kernels.cuh"

namespace sdr {

constexpr int TC_ROWS = 128;                 // A rows per tile (= TMEM lanes = threads)
constexpr int TC_OUT_PER_ROW = 16;
constexpr int TC_TILE_OUT = TC_ROWS * TC_OUT_PER_ROW;  // 2048 outputs per tile
constexpr int TC_D = 10;                     // decimation
constexpr int TC_Q = 16;                     // taps per phase (151 = 15*10 + 1)
constexpr int TC_FRONT = 16;                 // spare stream entries in front of row 0's window
constexpr int TC_STREAM = TC_FRONT + 16 * (TC_ROWS - 1) + 32;  // 2080 bytes per stream
constexpr int TC_NSTREAM = 2 * TC_D;         // 10 phases x {I,Q}
constexpr int TC_ND = 3;                     // signed base-256 digits per tap (23-bit fixed point)
constexpr int TC_N = 16 * TC_ND;             // 16 deltas x 3 digits = 48 accumulator columns
constexpr int TC_BP = TC_N * 32;             // bytes of one phase's B tile
constexpr int TC_HIST = 320;                 // raw history pairs the first tile reaches back
constexpr size_t TC_SMEM = (size_t)TC_NSTREAM * TC_STREAM + (size_t)TC_D * TC_BP;

struct RfTcArgs {
  RfArgs a;
  const int8_t *bmat;   // [10 phases][64 x 32] in canonical no-swizzle K-major core-matrix order
  const int32_t *hq;    // fixed-point taps, 160 entries (zero padded)
  long long corr;       // 128 * sum(hq): offset of the unsigned samples
  float scale;          // 2^-(S+7)
  int tiles_per_seg;
};

__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}

// 16 accumulator columns of this warp's 32 TMEM lanes.
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

// Recombine the three base-256 digit sums into the fixed-point FIR output and round once.
// v = d0 + 2^8 d1 + 2^16 d2 - corr is an integer below 2^39.  It is cut into H = v >> 17
// (|H| < 2^22) and L = v & (2^17-1): both are exactly representable floats, produced without a
// convert instruction by planting them in the mantissa of a magic constant, so
// fma(H, 2^17 * sc, L * sc) performs the ONLY rounding, and it is the round-to-nearest of v * sc
// (sc is a power of two).  (profiles/r1d, r1f: I2F.S64 was the most stalled instruction of the
// first version; a double-precision form throttled the FP64 pipe.)
struct TcScale {
  long long corr;
  float sc;     // 2^-(S+7)
  float sc17;   // 2^17 * sc
};
// 24 accumulator columns (8 outputs x 3 digits): one x16 and one x8 load.
__device__ __forceinline__ void tc_ld24(uint32_t taddr, uint32_t (&v)[24]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                 "=r"(v[23])
               : "r"(taddr + 16));
}

__device__ __forceinline__ float tc_combine(const uint32_t *d, const TcScale &k) {
  const int lo = (int32_t)d[0] + 256 * (int32_t)d[1];
  const long long v = (long long)(int32_t)d[2] * 65536 + lo - k.corr;
  const int H = (int)(v >> 17);
  const uint32_t L = (uint32_t)v & 0x1ffffu;
  const float Hf = __fsub_rn(__int_as_float(0x4B400000 + H), 12582912.0f);   // 2^23 + 2^22 + H
  const float Lf = __fsub_rn(__uint_as_float(0x4B000000u | L), 8388608.0f);
  return __fmaf_rn(Hf, k.sc17, __fmul_rn(Lf, k.sc));
}
// fmDemod on the fast path: same formula, approximate reciprocal (the fast variant is held to
// 100 dB / +-1 LSB against the reference, not to bit equality; I/Q already differ by ~1e-7).
__device__ __forceinline__ float tc_demod(float i, float q, float pi, float pq) {
  const float den = i * i + q * q;
  if (den == 0.0f) return 0.0f;
  return __fdividef(i * (q - pq) - q * (i - pi), den);
}

static __global__ void __launch_bounds__(TC_ROWS)
k_rf_demod_tc(const RfTcArgs g) {
  const RfArgs &a = g.a;
  extern __shared__ __align__(128) uint8_t tc_smem[];
  uint8_t *streams = tc_smem;                                      // [20][TC_STREAM]
  int8_t *bs = reinterpret_cast<int8_t *>(tc_smem + TC_NSTREAM * TC_STREAM);  // [10][TC_BP]
  __shared__ float carry_iq[2];
  __shared__ float edge[2][4];
  __shared__ long long red[2][4];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_slot;

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int b = blockIdx.y;
  const int n_tiles = (a.n_if + TC_TILE_OUT - 1) / TC_TILE_OUT;
  const int tile_begin = blockIdx.x * g.tiles_per_seg;
  const int tile_end = min(tile_begin + g.tiles_per_seg, n_tiles);
  if (tile_begin >= tile_end) return;
  const uint8_t *row = a.iq + (size_t)b * a.iq_stride;
  const uint8_t *hrow = a.hist + (size_t)b * 2 * a.rf_hist_len;
  const bool row_aligned = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);

  // ---- one-time setup: filter tiles, barrier, tensor memory ----
  for (int i = t; i < TC_D * TC_BP / 16; i += TC_ROWS)
    reinterpret_cast<uint4 *>(bs)[i] = __ldg(reinterpret_cast<const uint4 *>(g.bmat) + i);
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(tc_smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  // c S32, a U8, b S8, both K-major, N>>3, M>>4
  const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) |
                         ((uint32_t)(TC_ROWS >> 4) << 24);
  uint32_t phase = 0;
  const TcScale ks{g.corr, g.scale, g.scale * 131072.0f};

  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const long long j0 = (long long)tile * TC_TILE_OUT;
    // ---- 1. transpose raw bytes into the 20 phase streams ----
    // stream entry s' holds x[10*(j0 - 31 + s') - p]; a thread builds 4 consecutive entries
    // of all 20 streams from 40 consecutive input pairs (80 bytes, 2 bytes into an aligned
    // 96-byte window).
    for (int grp = t; grp < TC_STREAM / 4; grp += TC_ROWS) {
      const long long c_lo = 10 * j0 + 40ll * grp - 319;
      const long long w0 = 2 * c_lo - 2;  // stream byte of the window start (multiple of 16)
      if (row_aligned && w0 >= 0 && w0 + 96 <= 2 * a.n_rf) {
        uint32_t w[24];
        const uint4 *src = reinterpret_cast<const uint4 *>(row + w0);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const uint4 v = __ldg(src + k);
          w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
        }
#pragma unroll
        for (int p = 0; p < TC_D; ++p) {
#pragma unroll
          for (int comp = 0; comp < 2; ++comp) {
            // byte for entry ds: pair 9 + 10*ds - p of the 40, +2 bytes window offset
            const int P0 = 2 * (9 - p) + comp + 2, P1 = P0 + 20, P2 = P0 + 40, P3 = P0 + 60;
            const uint32_t ab = __byte_perm(w[P0 >> 2], w[P1 >> 2], (P0 & 3) | ((4 + (P1 & 3)) << 4));
            const uint32_t cd = __byte_perm(w[P2 >> 2], w[P3 >> 2], (P2 & 3) | ((4 + (P3 & 3)) << 4));
            *reinterpret_cast<uint32_t *>(streams + (2 * p + comp) * TC_STREAM + 4 * grp) =
                __byte_perm(ab, cd, 0x5410);
          }
        }
      } else {
        // edges of the capture (history before it, nothing after it) and unaligned rows
#pragma unroll 1
        for (int sc = 0; sc < TC_NSTREAM; ++sc) {
          const int p = sc >> 1, comp = sc & 1;
          uint32_t word = 0;
#pragma unroll 1
          for (int ds = 0; ds < 4; ++ds) {
            const long long pos = w0 + 2 * (9 + 10 * ds - p) + comp + 2;
            uint32_t val = 128;  // centred zero beyond either end
            if (pos < 0) {
              const long long h = 2ll * a.rf_hist_len + pos;
              if (h >= 0) val = hrow[h];
            } else if (pos < 2 * a.n_rf) {
              val = row[pos];
            }
            word |= val << (8 * ds);
          }
          *reinterpret_cast<uint32_t *>(streams + sc * TC_STREAM + 4 * grp) = word;
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;");  // generic-proxy writes -> tensor-core reads
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    // ---- 2. 20 MMAs: D_I, D_Q [128 x 64] += A_p(hankel bytes) * B_p(filter digits) ----
    if (t == 0) {
      const uint32_t s0 = tc_smem_u32(streams) + TC_FRONT, b0 = tc_smem_u32(bs);
#pragma unroll
      for (int comp = 0; comp < 2; ++comp) {
#pragma unroll
        for (int p = 0; p < TC_D; ++p) {
          const uint64_t da = tc_desc(s0 + (2 * p + comp) * TC_STREAM, 16, 128);
          const uint64_t db = tc_desc(b0 + p * TC_BP, (TC_N / 8) * 128, 128);
          const uint32_t acc = p > 0;
          asm volatile(
              "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem + comp * TC_N),
              "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          tc_smem_u32(&mbar)));
    }
    // ---- predecessor of the segment's first output ----
    if (tile == tile_begin) {
      if (j0 == 0) {
        if (t < 2) carry_iq[t] = a.prev_in[2 * b + t];
      } else {
        // output j0-1 evaluated in integers on the CUDA cores (same fixed-point taps; integer
        // sums are order independent): tap 10q+p meets xp[p][j0-1-q] = stream entry 30-q
        long long si = 0, sq = 0;
        for (int n = t; n < TC_D * TC_Q; n += TC_ROWS) {
          const int q = n / TC_D, p = n - q * TC_D;
          const long long h = g.hq[n];
          si += h * ((int)streams[(2 * p) * TC_STREAM + 30 - q] - 128);
          sq += h * ((int)streams[(2 * p + 1) * TC_STREAM + 30 - q] - 128);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          si += __shfl_xor_sync(0xffffffffu, si, o);
          sq += __shfl_xor_sync(0xffffffffu, sq, o);
        }
        if (lane == 0) {
          red[0][warp] = si;
          red[1][warp] = sq;
        }
      }
    }
    // ---- 3. wait for the accumulators ----
    {
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(tc_smem_u32(&mbar)), "r"(phase));
      }
      phase ^= 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    // ---- 4. epilogue: row t owns outputs j0 + 16t + delta ----
    float fi[TC_OUT_PER_ROW], fq[TC_OUT_PER_ROW];
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c = 0; c < 2; ++c) {   // 24 columns = 8 outputs per pass
      uint32_t vi[24], vq[24];
      tc_ld24(trow + 24 * c, vi);
      tc_ld24(trow + TC_N + 24 * c, vq);
      asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        fi[8 * c + k] = tc_combine(&vi[3 * k], ks);
        fq[8 * c + k] = tc_combine(&vq[3 * k], ks);
      }
    }
    // one-sample state: previous row's last output, by shuffle inside the warp and through
    // shared memory across warps / tiles
    float pi = __shfl_up_sync(0xffffffffu, fi[15], 1), pq = __shfl_up_sync(0xffffffffu, fq[15], 1);
    if (lane == 31) {
      edge[0][warp] = fi[15];
      edge[1][warp] = fq[15];
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (lane == 0) {
      if (warp > 0) {
        pi = edge[0][warp - 1];
        pq = edge[1][warp - 1];
      } else if (tile == tile_begin && j0 != 0) {
        pi = xmul(__ll2float_rn(red[0][0] + red[0][1] + red[0][2] + red[0][3]), g.scale);
        pq = xmul(__ll2float_rn(red[1][0] + red[1][1] + red[1][2] + red[1][3]), g.scale);
      } else {
        pi = carry_iq[0];
        pq = carry_iq[1];
      }
    }
    const long long jrow = j0 + 16 * t;
    float dm[TC_OUT_PER_ROW];
#pragma unroll
    for (int k = 0; k < TC_OUT_PER_ROW; ++k) {
      dm[k] = fm_demod_one(fi[k], fq[k], pi, pq);
      pi = fi[k];
      pq = fq[k];
    }
    float *drow = a.demod + (size_t)b * a.demod_stride + a.demod_off;
    if (jrow + TC_OUT_PER_ROW <= a.n_if && ((a.demod_stride | a.demod_off) & 3) == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        *reinterpret_cast<float4 *>(drow + jrow + 4 * k) = make_float4(dm[4 * k], dm[4 * k + 1], dm[4 * k + 2], dm[4 * k + 3]);
    } else {
#pragma unroll
      for (int k = 0; k < TC_OUT_PER_ROW; ++k)
        if (jrow + k < a.n_if) drow[jrow + k] = dm[k];
    }
    if (a.i_filt) {
#pragma unroll
      for (int k = 0; k < TC_OUT_PER_ROW; ++k)
        if (jrow + k < a.n_if) {
          a.i_filt[(size_t)b * a.tap_stride + jrow + k] = fi[k];
          a.q_filt[(size_t)b * a.tap_stride + jrow + k] = fq[k];
        }
    }
    {
      const long long last = (long long)a.n_if - 1 - jrow;  // position of the call's last output
      if (last >= 0 && last < TC_OUT_PER_ROW) {
#pragma unroll
        for (int k = 0; k < TC_OUT_PER_ROW; ++k)
          if (k == last) {
            a.prev_out[2 * b] = fi[k];
            a.prev_out[2 * b + 1] = fq[k];
          }
      }
    }
    __syncthreads();  // edge/carry reads done
    if (t == TC_ROWS - 1) {
      carry_iq[0] = fi[15];
      carry_iq[1] = fq[15];
    }
    // (the next tile's first __syncthreads orders this write before any read)
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}


// ---------------------------------------------------------------------------
// Throughput form of the same computation (the default).  The single-role kernel above is
// latency bound (profiles/r1b: long-scoreboard 44 %, 3 CTAs per SM, every tile waits for its
// global loads, then its MMAs, then runs its epilogue).  A first warp-specialised version
// (4 producer + 4 consumer warps, profiles/r1c) was slower: the producers stayed bound by
// global-load latency and the consumers starved.  This version keeps every warp busy with
// both kinds of work instead:
//   * the raw bytes of the NEXT tile are fetched with fully coalesced 16-byte cp.async copies
//     into a staging buffer while the current tile is processed (the transposer's own access
//     pattern -- 96-byte windows 80 bytes apart -- costs ~20 L1 lines per warp load when done
//     straight from global memory; from shared memory it is conflict free);
//   * 8 warps; all of them transpose (LDS.128 x6 + PRMT + STS per 40 input pairs);
//   * accumulators are double buffered: the MMAs of tile i run while the CTA does the
//     epilogue of tile i-1;
//   * a TMEM lane quarter is readable by warps w and w+4, so the two warps split a row's 16
//     outputs (delta 0-7 / 8-15); the one-sample demod state crosses via shared memory.
// ---------------------------------------------------------------------------
constexpr int TC_RAW_CHUNKS = (80 * (TC_STREAM / 4 - 1) + 96) / 16;   // 2601 16-byte chunks per tile
constexpr int TC_RAW = ((TC_RAW_CHUNKS * 16 + 127) / 128) * 128;     // raw staging bytes (padded)
constexpr size_t TC3_SMEM = (size_t)TC_RAW + (size_t)TC_NSTREAM * TC_STREAM + (size_t)TC_D * TC_BP;

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
      tc_smem_u32(bar)) : "memory");
}

// Transposes 40 input pairs (aligned 96-byte window, 2 bytes in) into 4 entries of all 20 streams.
__device__ __forceinline__ void tc_transpose_group(const uint32_t (&w)[24], uint8_t *streams, int grp) {
#pragma unroll
  for (int p = 0; p < TC_D; ++p) {
#pragma unroll
    for (int comp = 0; comp < 2; ++comp) {
      const int P0 = 2 * (9 - p) + comp + 2, P1 = P0 + 20, P2 = P0 + 40, P3 = P0 + 60;
      const uint32_t ab = __byte_perm(w[P0 >> 2], w[P1 >> 2], (P0 & 3) | ((4 + (P1 & 3)) << 4));
      const uint32_t cd = __byte_perm(w[P2 >> 2], w[P3 >> 2], (P2 & 3) | ((4 + (P3 & 3)) << 4));
      *reinterpret_cast<uint32_t *>(streams + (2 * p + comp) * TC_STREAM + 4 * grp) = __byte_perm(ab, cd, 0x5410);
    }
  }
}

static __global__ void __launch_bounds__(2 * TC_ROWS, 2)
k_rf_demod_tc3(const RfTcArgs g) {
  const RfArgs &a = g.a;
  extern __shared__ __align__(128) uint8_t tc_smem[];
  uint8_t *raw = tc_smem;                                   // [TC_RAW] staged input bytes
  uint8_t *streams = tc_smem + TC_RAW;                      // [20][TC_STREAM]
  int8_t *bs = reinterpret_cast<int8_t *>(tc_smem + TC_RAW + TC_NSTREAM * TC_STREAM);
  __shared__ float last_i[2][TC_ROWS], last_q[2][TC_ROWS];  // [half][row]: I,Q of delta 7 / 15
  __shared__ float carry_iq[2];
  __shared__ long long red[2][8];
  __shared__ __align__(8) uint64_t mma_done[2], raw_full;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rowi = (warp & 3) * 32 + lane;  // TMEM lane = A row served by this thread
  const int half = warp >> 2;               // which 8 of the row's 16 outputs
  const int b = blockIdx.y;
  const int n_tiles = (a.n_if + TC_TILE_OUT - 1) / TC_TILE_OUT;
  const int tile_begin = blockIdx.x * g.tiles_per_seg;
  const int tile_end = min(tile_begin + g.tiles_per_seg, n_tiles);
  if (tile_begin >= tile_end) return;
  const uint8_t *row = a.iq + (size_t)b * a.iq_stride;
  const uint8_t *hrow = a.hist + (size_t)b * 2 * a.rf_hist_len;
  const bool row_aligned = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);

  for (int i = tid; i < TC_D * TC_BP / 16; i += 2 * TC_ROWS)
    reinterpret_cast<uint4 *>(bs)[i] = __ldg(reinterpret_cast<const uint4 *>(g.bmat) + i);
  if (tid == 0) {
    mbar_init(&mma_done[0], 1);
    mbar_init(&mma_done[1], 1);
    mbar_init(&raw_full, 1);
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
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) |
                         ((uint32_t)(TC_ROWS >> 4) << 24);
  bool have_pred = false;  // predecessor of the segment's first output comes from `red`
  const TcScale ks{g.corr, g.scale, g.scale * 131072.0f};

  // Epilogue of tile `tile` (accumulator set buf, completion number `use` of mma_done[buf]).
  auto epilogue = [&](int tile, int buf, int use) {
    const long long j0 = (long long)tile * TC_TILE_OUT;
    mbar_wait(&mma_done[buf], use & 1);
    asm volatile("tcgen05.fence::after_thread_sync;");
    float fi[8], fq[8];
    const uint32_t trow = tmem + buf * 2 * TC_N + 24 * half + ((uint32_t)((warp & 3) * 32) << 16);
    {
      uint32_t vi[24], vq[24];
      tc_ld24(trow, vi);
      tc_ld24(trow + TC_N, vq);
      asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        fi[k] = tc_combine(&vi[3 * k], ks);
        fq[k] = tc_combine(&vq[3 * k], ks);
      }
    }
    last_i[half][rowi] = fi[7];
    last_q[half][rowi] = fq[7];
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    float pi, pq;
    if (half == 1) {            // delta 8 follows delta 7 of the same row
      pi = last_i[0][rowi];
      pq = last_q[0][rowi];
    } else if (rowi > 0) {      // delta 0 follows delta 15 of the previous row
      pi = last_i[1][rowi - 1];
      pq = last_q[1][rowi - 1];
    } else if (tile == tile_begin && have_pred) {
      pi = xmul(__ll2float_rn(red[0][0] + red[0][1] + red[0][2] + red[0][3] + red[0][4] + red[0][5] + red[0][6] + red[0][7]), g.scale);
      pq = xmul(__ll2float_rn(red[1][0] + red[1][1] + red[1][2] + red[1][3] + red[1][4] + red[1][5] + red[1][6] + red[1][7]), g.scale);
    } else {
      pi = carry_iq[0];
      pq = carry_iq[1];
    }
    const long long jrow = j0 + 16 * rowi + 8 * half;
    float dm[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      dm[k] = tc_demod(fi[k], fq[k], pi, pq);
      pi = fi[k];
      pq = fq[k];
    }
    float *drow = a.demod + (size_t)b * a.demod_stride + a.demod_off;
    if (jrow + 8 <= a.n_if && ((a.demod_stride | a.demod_off) & 3) == 0) {
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
    __syncthreads();  // last_* / carry reads done before they are rewritten
    if (half == 1 && rowi == TC_ROWS - 1) {
      carry_iq[0] = fi[7];
      carry_iq[1] = fq[7];
    }
  };

  if (tile_begin == 0 && tid < 2) carry_iq[tid] = a.prev_in[2 * b + tid];

  // Stage the raw bytes of a tile: stream bytes [20*j0 - 640, +TC_RAW_CHUNKS*16).  A tile that
  // lies entirely inside the capture is one bulk asynchronous copy (TMA, completes on raw_full);
  // tiles that touch the history in front of the capture or its end are assembled chunk by chunk.
  auto tile_is_bulk = [&](int tile) -> bool {
    const long long wbase = 20ll * tile * TC_TILE_OUT - 640;
    return row_aligned && wbase >= 0 && wbase + 16ll * TC_RAW_CHUNKS <= 2 * a.n_rf;
  };
  uint32_t raw_phase = 0;
  auto issue_raw = [&](int tile) {
    const long long wbase = 20ll * tile * TC_TILE_OUT - 640;
    if (tile_is_bulk(tile)) {
      if (tid == 0) {
        constexpr uint32_t BYTES = TC_RAW_CHUNKS * 16;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(&raw_full)), "r"(BYTES)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                tc_smem_u32(raw)),
            "l"(row + wbase), "r"(BYTES), "r"(tc_smem_u32(&raw_full))
            : "memory");
      }
      return;
    }
    for (int q = tid; q < TC_RAW_CHUNKS; q += 2 * TC_ROWS) {
      const long long pos = wbase + 16ll * q;
      if (row_aligned && pos >= 0 && pos + 16 <= 2 * a.n_rf) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc_smem_u32(raw + 16 * q)), "l"(row + pos)
                     : "memory");
      } else {
        for (int k = 0; k < 16; ++k) {
          const long long p = pos + k;
          uint8_t val = 128;  // centred zero beyond either end
          if (p < 0) {
            const long long h = 2ll * a.rf_hist_len + p;
            if (h >= 0) val = hrow[h];
          } else if (p < 2 * a.n_rf) {
            val = row[p];
          }
          raw[16 * q + k] = val;
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto wait_raw = [&](int tile) {
    if (tile_is_bulk(tile)) {
      mbar_wait(&raw_full, raw_phase);
      raw_phase ^= 1;
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
  };
  issue_raw(tile_begin);

  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int it = tile - tile_begin, buf = it & 1;
    wait_raw(tile);                                                      // this tile's bytes have landed
    if (it > 0) mbar_wait(&mma_done[buf ^ 1], ((it - 1) >> 1) & 1);      // MMAs of tile it-1 have read `streams`
    __syncthreads();                                                     // everyone's copies are visible
    // ---- 1. transpose raw -> 20 phase streams (each thread: 40 input pairs per group) ----
    for (int grp = tid; grp < TC_STREAM / 4; grp += 2 * TC_ROWS) {
      uint32_t w[24];
      const uint4 *src = reinterpret_cast<const uint4 *>(raw + 80 * grp);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const uint4 v = src[k];
        w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
      }
      tc_transpose_group(w, streams, grp);
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    // ---- 2. MMAs of this tile into accumulator set buf (drained by the epilogue of tile it-2) ----
    if (tid == 0) {
      const uint32_t s0 = tc_smem_u32(streams) + TC_FRONT, b0 = tc_smem_u32(bs);
      const uint32_t d0 = tmem + buf * 2 * TC_N;
#pragma unroll
      for (int comp = 0; comp < 2; ++comp) {
#pragma unroll
        for (int p = 0; p < TC_D; ++p) {
          const uint64_t da = tc_desc(s0 + (2 * p + comp) * TC_STREAM, 16, 128);
          const uint64_t db = tc_desc(b0 + p * TC_BP, (TC_N / 8) * 128, 128);
          const uint32_t acc = p > 0;
          asm volatile(
              "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d0 + comp * TC_N),
              "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
      }
      tc_commit(&mma_done[buf]);
    }
    // ---- predecessor of the segment's first output, in integers (see the single-role kernel) ----
    if (tile == tile_begin && tile != 0) {
      long long si = 0, sq = 0;
      for (int n = tid; n < TC_D * TC_Q; n += 2 * TC_ROWS) {
        const int q = n / TC_D, p = n - q * TC_D;
        const long long h = g.hq[n];
        si += h * ((int)streams[(2 * p) * TC_STREAM + 30 - q] - 128);
        sq += h * ((int)streams[(2 * p + 1) * TC_STREAM + 30 - q] - 128);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        si += __shfl_xor_sync(0xffffffffu, si, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
      }
      if (lane == 0) {
        red[0][warp] = si;
        red[1][warp] = sq;
      }
      have_pred = true;  // visible to everyone after the epilogue's first __syncthreads
    }
    // ---- 3. next tile's bytes start flowing; epilogue of the previous tile meanwhile ----
    // `raw` is free: every thread's reads precede the __syncthreads above (and that barrier's
    // fence.proxy.async orders them before the bulk copy's async-proxy writes)
    if (tile + 1 < tile_end) issue_raw(tile + 1);
    if (it > 0) epilogue(tile - 1, buf ^ 1, (it - 1) >> 1);
  }
  {
    const int it = tile_end - 1 - tile_begin;
    epilogue(tile_end - 1, it & 1, it >> 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}


// Tried and dropped (profiles/r1g): one CTA per SM with 16 warps sharing each tile and the copy of
// tile i+2, the transpose and MMAs of tile i and the epilogue of tile i-1 all overlapped.  It was
// slower (0.84 ms vs 0.56 ms): the per-thread fixed costs are paid by 512 threads per tile (7450
// warp instructions per tile instead of 5500) and 38 % of the stall samples sat on block barriers.

}  // namespace sdr
