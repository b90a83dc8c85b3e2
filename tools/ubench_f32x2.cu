// Micro-benchmark: issue rate of scalar FMUL+FADD vs packed mul.f32x2 + add.f32x2 on sm_100a.
// Decides whether the exact (separately rounded) FIR inner loop can be halved in issue slots.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_scalar(float *out, float a, float b, int iters) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __fadd_rn(x[i], __fmul_rn(x[(i + 1) & 7], a));
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + b;
}
__global__ void k_packed(float *out, float a, float b, int iters) {
  unsigned long long x[4];
  for (int i = 0; i < 4; ++i) {
    float lo = threadIdx.x * 0.001f + 2 * i, hi = lo + 1;
    asm("mov.b64 %0, {%1,%2};" : "=l"(x[i]) : "f"(lo), "f"(hi));
  }
  unsigned long long aa;
  asm("mov.b64 %0, {%1,%1};" : "=l"(aa) : "f"(a));
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      unsigned long long p;
      asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(p) : "l"(x[(i + 1) & 3]), "l"(aa));
      asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(x[i]) : "l"(x[i]), "l"(p));
    }
  }
  float s = 0;
  for (int i = 0; i < 4; ++i) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + b;
}
int main() {
  float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    float ms;
    cudaEventRecord(e0); k_scalar<<<148 * 8, 256>>>(d, 1.0001f, 0.f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double macs = 148.0 * 8 * 256 * iters * 8;
    printf("scalar FMUL+FADD : %.3f ms  %.2f TMAC/s\n", ms, macs / ms * 1e-9);
    cudaEventRecord(e0); k_packed<<<148 * 8, 256>>>(d, 1.0001f, 0.f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("packed mul/add x2: %.3f ms  %.2f TMAC/s\n", ms, macs / ms * 1e-9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
