// Dependent-chain latency of the pieces of the exact stereo PLL step (csrc/kernels.cuh k_pll),
// one warp, each piece fed back into itself.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I software-defined-radio_b200/csrc \
//        -o tools/ubench_pll_chain.bin tools/ubench_pll_chain.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "libm_exact.cuh"

using namespace sdr;
namespace sdr {
void set_error(const std::string &) {}
int cuda_fail(cudaError_t, const char *, const char *, int) { return -1; }
}  // namespace sdr

template <int OP>
__global__ void chain(float *out, float a, float b, int iters) {
  float x = a + 0.001f * threadIdx.x, y = b;
  float fbI = 0.8f, fbQ = 0.6f, integ = 0.0f, phase = 0.1f, trig = 1000.0f;
  const double w = 0.4974188368183839;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) {  // atan2f only
      float e;
      if (!atan2f_common(x, y, e)) e = atan2f_glibc(x, y);
      x = xadd(e, 0.3f);
    }
    if (OP == 1) {  // sincosf (large argument path) only
      float s, c;
      sincosf_glibc_bf(x, s, c);
      x = xadd(xmul(s, 100.0f), 500.0f);
      y = c;
    }
    if (OP == 2) {  // IEEE division + one add (bounded orbit)
      x = xadd(xdiv(1.7f, x), 0.3f);
    }
    if (OP == 3) {  // loop filter + trigArg (float -> double -> float)
      integ = xadd(integ, xmul(1e-4f, x));
      phase = xadd(xadd(phase, xmul(0.02f, x)), integ);
      trig = xadd(trig, 1.0f);
      x = __double2float_rn(__dadd_rn(__dmul_rn(w, (double)trig), (double)phase));
      x = xmul(x, 1e-3f);
    }
    if (OP == 4) {  // the whole step as in k_pll (without the NCO output cosine)
      const float eI = xmul(x, fbI), eQ = xmul(x, -fbQ);
      float e;
      if (!atan2f_common(eQ, eI, e)) e = atan2f_glibc(eQ, eI);
      integ = xadd(integ, xmul(3.555e-4f, e));
      phase = xadd(xadd(phase, xmul(0.02666f, e)), integ);
      trig = xadd(trig, 1.0f);
      const float ta = __double2float_rn(__dadd_rn(__dmul_rn(w, (double)trig), (double)phase));
      sincosf_glibc_bf(ta, fbQ, fbI);
    }
    if (OP == 5) {  // whole step + NCO cosine (off the chain, same thread)
      const float eI = xmul(x, fbI), eQ = xmul(x, -fbQ);
      float e;
      if (!atan2f_common(eQ, eI, e)) e = atan2f_glibc(eQ, eI);
      integ = xadd(integ, xmul(3.555e-4f, e));
      phase = xadd(xadd(phase, xmul(0.02666f, e)), integ);
      trig = xadd(trig, 1.0f);
      const float ta = __double2float_rn(__dadd_rn(__dmul_rn(w, (double)trig), (double)phase));
      sincosf_glibc_bf(ta, fbQ, fbI);
      y = xadd(y, cosf_glibc_bf(xadd(xmul(ta, 2.0f), 0.0f)));
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x + y + fbI + fbQ + integ + phase;
  if (threadIdx.x == 0) out[64] = (float)(t1 - t0) / (float)iters;
}

int main() {
  float *d;
  cudaMalloc(&d, 128 * sizeof(float));


  const char *names[] = {"atan2f", "sincosf (large arg)", "IEEE division", "loop filter + trigArg", "whole step",
                         "whole step + NCO cos"};
  for (int op = 0; op < 6; ++op) {
    for (int rep = 0; rep < 2; ++rep) {
      switch (op) {
        case 0: chain<0><<<1, 32>>>(d, 0.7f, 0.9f, 20000); break;
        case 1: chain<1><<<1, 32>>>(d, 700.0f, 0.9f, 20000); break;
        case 2: chain<2><<<1, 32>>>(d, 0.7f, 0.9f, 20000); break;
        case 3: chain<3><<<1, 32>>>(d, 0.7f, 0.9f, 20000); break;
        case 4: chain<4><<<1, 32>>>(d, 0.05f, 0.9f, 20000); break;
        case 5: chain<5><<<1, 32>>>(d, 0.05f, 0.9f, 20000); break;
      }
      cudaDeviceSynchronize();
    }
    float cyc;
    cudaMemcpy(&cyc, d + 64, sizeof(float), cudaMemcpyDeviceToHost);
    printf("%-24s %7.1f cycles per iteration\n", names[op], cyc);
  }
  return 0;
}
