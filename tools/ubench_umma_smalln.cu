// Micro-benchmark: cost of small-N tcgen05.mma (kind::f16, M = 128, K = 16, operands in shared memory)
// as a function of N and of how many INDEPENDENT accumulators the MMAs rotate over.  Answers whether
// a stream of short MMAs into one accumulator is bound by a per-MMA latency (then rotating over
// several accumulators helps) or by operand fetch / issue (then it does not).
// One CTA, one issuing thread, 256 MMAs per measurement, cycles from issue of the first to completion
// of the last (tcgen05.commit -> mbarrier).   nvcc -arch=sm_100a -O2 -o tools/ubench_umma_smalln.bin ...
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}

__global__ void __launch_bounds__(128) k_bench(long long *out) {
  extern __shared__ __align__(128) uint8_t smem[];   // A: 8 chunks x 128 rows x 16 B = 16 KB, B: 8 chunks x 256 rows x 16 B = 32 KB
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_s;
  const int t = threadIdx.x;
  for (int i = t; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;  // 1.0h
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  if (t < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_s;
  if (t == 0) {
    int row = 0;
    uint32_t phase = 0;
    const int Ns[5] = {16, 32, 64, 128, 256};
    for (int ni = 0; ni < 5; ++ni) {
      const int N = Ns[ni];
      for (int chains = 1; chains <= 8 && chains * N <= 512; chains *= 2) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        long long t_issue = 0, t_done = 0;
        for (int rep = 0; rep < 2; ++rep) {
          const long long t0 = clock64();
          for (int i = 0; i < 256; ++i) {
            const uint64_t da = make_desc(smem_u32(smem) + (i & 3) * 4096, 2048, 128);
            const uint64_t db = make_desc(smem_u32(smem) + 16384 + (i & 3) * 8192, 4096, 128);
            const uint32_t d = tmem + (i % chains) * N;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                         "l"(da), "l"(db), "r"(idesc), "r"(i >= chains ? 1u : 0u));
          }
          const long long t1 = clock64();
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)));
          uint32_t done = 0;
          while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(phase));
          phase ^= 1;
          const long long t2 = clock64();
          t_issue = t1 - t0;
          t_done = t2 - t0;
        }
        out[4 * row] = N; out[4 * row + 1] = chains; out[4 * row + 2] = t_issue; out[4 * row + 3] = t_done;
        ++row;
      }
    }
    out[4 * row] = -1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
  long long *d, h[4 * 64];
  cudaMalloc(&d, sizeof h);
  cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
  k_bench<<<1, 128, 49152>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  printf("M=128 K=16 f16, 256 MMAs:   N  chains  issue_cycles/MMA  total_cycles/MMA\n");
  for (int r = 0; h[4 * r] >= 0 && r < 63; ++r)
    printf("  %4lld  %3lld   %8.1f   %8.1f\n", h[4 * r], h[4 * r + 1], h[4 * r + 2] / 256.0, h[4 * r + 3] / 256.0);
  return 0;
}
