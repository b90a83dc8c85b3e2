// rf_tc.cuh -- tensor-core RF front end (SDR_VARIANT_FAST, mono, rf_decim = 10).
//
// Why: the bit-exact CUDA-core front end is FP32-issue bound (profiles/: FMA pipe is the
// top pipe, DRAM < 10 %): 30.2 multiply-adds per input sample against 2 bytes, and the
// exact form needs two instructions per multiply-add.  north_star admits a Toeplitz
// GEMM in exactly that situation.
//
// How: the raw interleaved uint8 stream is consumed by tcgen05.mma kind::i8 directly,
// with no conversion, de-interleave or im2col pass.  Row m of the A operand is the 320
// bytes starting 16 bytes (8 complex samples) after row m-1: a shared-memory matrix
// descriptor with leading-byte-offset 16 and stride-byte-offset 128 turns the byte
// stream into that overlapping-row (Hankel) matrix in place.  The B operand holds the
// filter: taps are scaled to 31-bit fixed point and split into four signed base-256
// digits; column (theta, I|Q, digit) has digit(h[t]) at byte 2*(theta+150-t) + (I:0,Q:1).
// Row m therefore yields, for the one output j with 10*j = c_m + theta + 150, the exact
// integer sums  sum_t digit_d(h[t]) * u8[...]  in int32.  The epilogue recombines the
// digits in int64, removes the 128 offset of the unsigned samples, and rounds ONCE to
// float: the result is the exactly-rounded fixed-point FIR output (tap quantisation
// 2^-34), which differs from the reference's sequential float sum only by the
// reference's own accumulated rounding (~1e-7 relative; >= 100 dB SNR, tests).
// fmDemod follows in the same kernel with the reference's float operations.
#pragma once

#include "kernels.cuh"

namespace sdr {

constexpr int TC_ROWS = 128;                // A rows per tile (TMEM lanes)
constexpr int TC_ROW_SHIFT = 8;             // complex samples between consecutive rows
constexpr int TC_TILE = TC_ROWS * TC_ROW_SHIFT;  // 1024 complex samples per tile
constexpr int TC_LEAD = 152;                // window of row 0 starts this far before the tile
constexpr int TC_EXTRA = 16;                // extra samples in front (predecessor output)
constexpr int TC_K = 320;                   // bytes per row window (10 k-steps of 32)
constexpr int TC_N = 32;                    // 4 thetas x (I,Q) x 4 digits
constexpr int TC_STAGE = 2 * TC_EXTRA + 16 * (TC_ROWS - 1) + TC_K;  // 2384 bytes
constexpr int TC_STAGE_PAD = 2432;
constexpr int TC_MAX_OUT = 104;             // outputs owned by one tile (<= 103)

struct RfTcArgs {
  RfArgs a;
  const int8_t *bmat;   // [TC_N x TC_K] in canonical no-swizzle K-major core-matrix order
  const int32_t *hq;    // fixed-point taps, 151 entries (zero padded)
  long long corr;       // 128 * sum(hq): offset of the unsigned samples
  float scale;          // 2^-(S+7)
  int ntaps;
  int tiles_per_seg;
};

__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}

static __global__ void __launch_bounds__(TC_ROWS)
k_rf_demod_tc(const RfTcArgs g) {
  const RfArgs &a = g.a;
  __shared__ __align__(128) uint8_t stage[TC_STAGE_PAD];
  __shared__ __align__(128) int8_t bs[TC_N * TC_K];
  __shared__ float iq_i[TC_MAX_OUT + 1], iq_q[TC_MAX_OUT + 1];
  __shared__ long long red[2][4];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_slot;

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int b = blockIdx.y;
  const long long n_tiles = (a.n_rf + TC_TILE - 1) / TC_TILE;
  const long long tile_begin = (long long)blockIdx.x * g.tiles_per_seg;
  const long long tile_end = min(tile_begin + (long long)g.tiles_per_seg, n_tiles);
  if (tile_begin >= tile_end) return;
  const uint8_t *row = a.iq + (size_t)b * a.iq_stride;
  const uint8_t *hrow = a.hist + (size_t)b * 2 * a.rf_hist_len;
  const bool row_aligned = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);

  // ---- one-time setup: filter matrix, barrier, tensor memory ----
  for (int i = t; i < TC_N * TC_K / 16; i += TC_ROWS)
    reinterpret_cast<uint4 *>(bs)[i] = __ldg(reinterpret_cast<const uint4 *>(g.bmat) + i);
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(tc_smem_u32(&tmem_slot)));
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

  for (long long tile = tile_begin; tile < tile_end; ++tile) {
    const long long tile_c0 = tile * TC_TILE;
    const long long byte0 = 2 * (tile_c0 - TC_LEAD - TC_EXTRA);  // stream byte of stage[0]
    // ---- stage the raw bytes (no conversion): 16-byte chunks ----
    for (int q = t; q < TC_STAGE_PAD / 16; q += TC_ROWS) {
      const long long pos = byte0 + 16ll * q;
      uint4 v;
      if (row_aligned && pos >= 0 && pos + 16 <= 2 * a.n_rf) {
        v = __ldg(reinterpret_cast<const uint4 *>(row + pos));
      } else {
        uint8_t tmp[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const long long p = pos + k;
          uint8_t val = 128;  // centred zero beyond either end
          if (p < 0) {
            const long long h = 2ll * a.rf_hist_len + p;
            if (h >= 0) val = hrow[h];
          } else if (p < 2 * a.n_rf) {
            val = row[p];
          }
          tmp[k] = val;
        }
        v = *reinterpret_cast<uint4 *>(tmp);
      }
      reinterpret_cast<uint4 *>(stage)[q] = v;
    }
    asm volatile("fence.proxy.async.shared::cta;");  // generic-proxy writes -> tensor-core reads
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    // ---- 10 MMAs: D[128 x 32] = A(hankel bytes) * B(filter digits) ----
    if (t == 0) {
      const uint32_t a0 = tc_smem_u32(stage) + 2 * TC_EXTRA, b0 = tc_smem_u32(bs);
#pragma unroll
      for (int ks = 0; ks < TC_K / 32; ++ks) {
        const uint64_t da = tc_desc(a0 + 32 * ks, 16, 128);
        const uint64_t db = tc_desc(b0 + ks * 2 * (TC_N / 8) * 128, (TC_N / 8) * 128, 128);
        const uint32_t acc = ks > 0;
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
            "l"(da), "l"(db), "r"(idesc), "r"(acc));
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          tc_smem_u32(&mbar)));
    }
    // ---- outputs owned by this tile; predecessor of the first one ----
    const long long base10 = tile_c0 - 2;                       // 10*j lies in [base10, base10+1024)
    const long long j_lo = (base10 + 9 >= 0) ? (base10 + 9) / 10 : 0;
    if (tile == tile_begin) {
      if (j_lo == 0) {
        if (t < 2) (t == 0 ? iq_i : iq_q)[0] = a.prev_in[2 * b + t];
      } else {
        // integer evaluation of output j_lo-1 on the CUDA cores (same fixed-point taps,
        // integer sums are order independent): newest sample 10*(j_lo-1)
        const long long newest = 10 * (j_lo - 1);
        long long si = 0, sq = 0;
        for (int n = t; n < g.ntaps; n += TC_ROWS) {
          const long long c = newest - n;
          const int off = (int)(2 * c - byte0);
          const long long h = g.hq[n];
          si += h * ((int)stage[off] - 128);
          sq += h * ((int)stage[off + 1] - 128);
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
    // ---- wait for the accumulator ----
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
    uint32_t v[32];
    {
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
          "%28,%29,%30,%31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
            "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
            "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;");
    }
    // row t: window starts at c = tile_c0 - 152 + 8t; it owns the output with
    // 10*j = c + 150 + theta, theta in {0,2,4,6} (rows with theta == 8 own none).
    // All per-thread index math is 32-bit, relative to 10*j_lo.
    const int u2 = (int)(base10 - 10 * j_lo) + 8 * t + 10;   // (c + 150) - 10*j_lo + 10  >= 1
    const int theta = (10 - u2 % 10) % 10;
    const int jrel = (u2 + theta) / 10 - 1;
    const long long j = j_lo + jrel;
    const bool owns = theta < 8 && j < a.n_if;
    const int th = theta >> 1;
    long long vi = 0, vq = 0;
    {
      int32_t di[4], dq[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        di[d] = (th == 0) ? v[d] : (th == 1) ? v[8 + d] : (th == 2) ? v[16 + d] : v[24 + d];
        dq[d] = (th == 0) ? v[4 + d] : (th == 1) ? v[12 + d] : (th == 2) ? v[20 + d] : v[28 + d];
      }
      vi = (((long long)di[3] * 256 + di[2]) * 256 + di[1]) * 256 + di[0] - g.corr;
      vq = (((long long)dq[3] * 256 + dq[2]) * 256 + dq[1]) * 256 + dq[0] - g.corr;
    }
    const float fi = xmul(__ll2float_rn(vi), g.scale);
    const float fq = xmul(__ll2float_rn(vq), g.scale);
    const int slot = jrel + 1;
    if (owns) {
      iq_i[slot] = fi;
      iq_q[slot] = fq;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (owns) {
      float pi = iq_i[slot - 1], pq = iq_q[slot - 1];
      if (slot == 1 && tile == tile_begin && j_lo > 0) {
        // predecessor of the segment's first output: the integer sums reduced above
        pi = xmul(__ll2float_rn(red[0][0] + red[0][1] + red[0][2] + red[0][3]), g.scale);
        pq = xmul(__ll2float_rn(red[1][0] + red[1][1] + red[1][2] + red[1][3]), g.scale);
      }
      const float d = fm_demod_one(fi, fq, pi, pq);
      a.demod[(size_t)b * a.demod_stride + a.demod_off + j] = d;
      if (a.i_filt) {
        a.i_filt[(size_t)b * a.tap_stride + j] = fi;
        a.q_filt[(size_t)b * a.tap_stride + j] = fq;
      }
      if (j == a.n_if - 1) {
        a.prev_out[2 * b] = fi;
        a.prev_out[2 * b + 1] = fq;
      }
    }
    // carry the tile's last output to slot 0 for the next tile
    const long long j_hi = min((long long)a.n_if, (base10 + 1024 + 9) / 10);  // exclusive
    __syncthreads();
    if (t < 2 && j_hi > j_lo) {
      float *arr = (t == 0 ? iq_i : iq_q);
      arr[0] = arr[(int)(j_hi - j_lo)];
    }
    // (the next iteration's first __syncthreads orders this write before any read)
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem));
}

}  // namespace sdr
