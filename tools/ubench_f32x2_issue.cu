// Micro-benchmark: does the packed form of the EXACT multiply-add (FMUL2 + FFMA2 with a run-time
// 1.0 multiplier: two roundings per lane, never contracted) free issue slots on sm_100a?
// Scalar FMUL + FADD is issue-bound (2 slots per multiply-add); the packed pair issues 2 instructions per
// 2 multiply-adds.  K extra ALU-pipe instructions (and optionally one LDS.128) per two multiply-adds stand in
// for the address / load overhead of a real FIR loop.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ void alu(unsigned &y, unsigned k) { asm volatile("lop3.b32 %0, %0, %1, %0, 0x96;" : "+r"(y) : "r"(k)); }

template <int K, int LDS>
__global__ void k_scalar(float *out, float a, int iters) {
  __shared__ float4 sm[256];
  sm[threadIdx.x] = make_float4(1.f, 2.f, 3.f, 4.f);
  __syncthreads();
  float x[8]; unsigned y = threadIdx.x; float4 v = make_float4(0, 0, 0, 0);
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      x[i] = __fadd_rn(x[i], __fmul_rn(x[(i + 2) & 7], a));
      x[i + 1] = __fadd_rn(x[i + 1], __fmul_rn(x[(i + 3) & 7], a));
#pragma unroll
      for (int k = 0; k < K; ++k) alu(y, it);
    }
    if (LDS) { float4 w = sm[(threadIdx.x + it) & 255]; v.x += w.x; }
  }
  float s = v.x; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)y;
}
template <int K, int LDS>
__global__ void k_packed(float *out, float a, float one, int iters) {
  __shared__ float4 sm[256];
  sm[threadIdx.x] = make_float4(1.f, 2.f, 3.f, 4.f);
  __syncthreads();
  u64 x[4]; unsigned y = threadIdx.x; float4 v = make_float4(0, 0, 0, 0);
  for (int i = 0; i < 4; ++i) x[i] = pk(threadIdx.x * 0.001f + 2 * i, threadIdx.x * 0.001f + 2 * i + 1);
  const u64 aa = pk(a, a), one2 = pk(one, one);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x[i] = fma2(mul2(x[(i + 1) & 3], aa), one2, x[i]);
#pragma unroll
      for (int k = 0; k < K; ++k) alu(y, it);
    }
    if (LDS) { float4 w = sm[(threadIdx.x + it) & 255]; v.x += w.x; }
  }
  float s = v.x;
  for (int i = 0; i < 4; ++i) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)y;
}
template <typename F> static void run(const char *name, F launch) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000; float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); launch(iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  const double macs = 148.0 * 8 * 256 * iters * 8;
  printf("%-40s %.3f ms  %.2f TMAC/s\n", name, ms, macs / ms * 1e-9);
}
int main() {
  float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
#define RUN(K, L) \
  run("scalar FMUL+FADD  K=" #K " LDS=" #L, [&](int it) { k_scalar<K, L><<<148 * 8, 256>>>(d, 1.0001f, it); }); \
  run("packed FMUL2+FFMA2 K=" #K " LDS=" #L, [&](int it) { k_packed<K, L><<<148 * 8, 256>>>(d, 1.0001f, 1.0f, it); });
  RUN(0, 0) RUN(1, 0) RUN(2, 0) RUN(3, 0) RUN(1, 1)
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
