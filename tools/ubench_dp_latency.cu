// Dependent-issue latency of the double-precision pipe on B200 (one warp, one chain):
// what bounds the per-sample step of the sequential PLL kernels (csrc/rds.cu k_rds_pll).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_dp_latency.bin tools/ubench_dp_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(double *out, double a, double b, int iters) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (OP == 0) x = fma(x, b, a);
      if (OP == 1) x = __dadd_rn(x, b);
      if (OP == 2) x = __dmul_rn(x, b);
      if (OP == 3) x = rint(x * b);                      // DMUL + FRND.F64
      if (OP == 4) x = __longlong_as_double(__double_as_longlong(x) ^ 0x10);  // integer op on a double
      if (OP == 5) x = (x > a) ? b : x + a;              // DSETP + select + DADD
      if (OP == 6) x = (__double2hiint(x) < 0) ? b : x + a;  // ISETP + select + DADD
      if (OP == 7) { float f = (float)x; f = fmaf(f, 1.0001f, 0.5f); x = (double)f; }  // F2F + FFMA + F2F
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) out[64] = (double)(t1 - t0) / (16.0 * iters);
}

int main() {
  double *d;
  cudaMalloc(&d, 128 * sizeof(double));
  const char *names[] = {"DFMA", "DADD", "DMUL", "DMUL+FRND.F64", "LOP on double", "DSETP+sel+DADD",
                         "ISETP(hi)+sel+DADD", "F2F+FFMA+F2F"};
  for (int op = 0; op < 8; ++op) {
    for (int rep = 0; rep < 2; ++rep) {
      switch (op) {
        case 0: chain<0><<<1, 32>>>(d, 1.0, 0.999, 4096); break;
        case 1: chain<1><<<1, 32>>>(d, 1.0, 0.999, 4096); break;
        case 2: chain<2><<<1, 32>>>(d, 1.0, 0.999, 4096); break;
        case 3: chain<3><<<1, 32>>>(d, 1.0, 1.3, 4096); break;
        case 4: chain<4><<<1, 32>>>(d, 1.0, 0.999, 4096); break;
        case 5: chain<5><<<1, 32>>>(d, 1.0, 0.999, 4096); break;
        case 6: chain<6><<<1, 32>>>(d, 1.0, 0.999, 4096); break;
        case 7: chain<7><<<1, 32>>>(d, 1.0, 0.999, 4096); break;
      }
      cudaDeviceSynchronize();
    }
    double cyc;
    cudaMemcpy(&cyc, d + 64, sizeof(double), cudaMemcpyDeviceToHost);
    printf("%-22s %6.1f cycles per link\n", names[op], cyc);
  }
  return 0;
}
