// Feasibility test for the tensor-core front end: tcgen05.mma kind::i8 with an A-operand shared
// memory descriptor whose leading-byte-offset (16 B) is smaller than a core matrix, so that the
// 128 rows are OVERLAPPING windows of one byte stream (row m starts 16 bytes after row m-1):
//   D[m][n] = sum_k x[16*m + k] * B[n][k],  x unsigned 8-bit, B signed 8-bit, D int32.
// That is a Hankel matrix expressed without copying any data.  Prints PASS/FAIL vs the host.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 32, KSTEPS = 10, K = 32 * KSTEPS;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // layout_type = 0 (no swizzle), base_offset = 0
}

__global__ void __launch_bounds__(128) k_test(const uint8_t *x, const int8_t *bmat, int32_t *out) {
  __shared__ __align__(128) uint8_t xs[16 * M + K + 64];
  __shared__ __align__(128) int8_t bs[N * K];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < 16 * M + K; i += 128) xs[i] = x[i];
  // canonical no-swizzle K-major: core matrix (8 rows x 16 bytes) index = q*(N/8) + nb
  for (int i = t; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    const int off = ((k / 16) * (N / 8) + n / 8) * 128 + (n % 8) * 16 + (k % 16);
    bs[off] = bmat[i];
  }
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  // make generic-proxy smem writes visible to the async (tensor core) proxy
  asm volatile("fence.proxy.async.shared::cta;");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  if (t == 0) {
    // c_format S32 (2), a_format U8 (0), b_format S8 (1), K-major both, N>>3, M>>4
    const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int ks = 0; ks < KSTEPS; ++ks) {
      const uint64_t da = make_desc(smem_u32(xs) + 32 * ks, 16, 128);              // overlapping rows
      const uint64_t db = make_desc(smem_u32(bs) + ks * 2 * (N / 8) * 128, (N / 8) * 128, 128);
      const uint32_t acc = ks > 0;
      asm volatile(
          "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(acc));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)));
  }
  // wait for the MMAs (phase 0)
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(done)
          : "r"(smem_u32(&mbar)), "r"(0));
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t v[32];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  for (int n = 0; n < N; ++n) out[t * N + n] = (int32_t)v[n];
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem));
}

int main() {
  std::vector<uint8_t> x(16 * M + K);
  std::vector<int8_t> b(N * K);
  srand(1);
  for (auto &v : x) v = rand() & 255;
  for (auto &v : b) v = (int8_t)(rand() & 255);
  uint8_t *dx; int8_t *db; int32_t *dout;
  cudaMalloc(&dx, x.size()); cudaMalloc(&db, b.size()); cudaMalloc(&dout, M * N * 4);
  cudaMemcpy(dx, x.data(), x.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size(), cudaMemcpyHostToDevice);
  cudaMemset(dout, 0xff, M * N * 4);
  k_test<<<1, 128>>>(dx, db, dout);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<int32_t> out(M * N);
  cudaMemcpy(out.data(), dout, M * N * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      long long s = 0;
      for (int k = 0; k < K; ++k) s += (long long)x[16 * m + k] * b[n * K + k];
      if ((int32_t)s != out[m * N + n]) {
        if (bad < 8) printf("mismatch m=%d n=%d want %lld got %d\n", m, n, s, out[m * N + n]);
        ++bad;
      }
    }
  printf("%s (%d mismatches of %d)\n", bad ? "FAIL" : "PASS", bad, M * N);
  return bad ? 1 : 0;
}
