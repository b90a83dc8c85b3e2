// common.cuh -- shared device/host helpers for the B200 FM receiver kernels.
//
// Arithmetic contract: the reference (src/filter.cpp) is built for baseline
// x86-64, so every float multiply and add rounds separately (mulss/addss, no
// FMA), division is IEEE and denormals are kept.  The helpers below spell that
// out with round-to-nearest intrinsics so that no compiler flag can contract
// them; the library is additionally compiled with -fmad=false.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

namespace sdr {

__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
// acc + h*x with two roundings: one step of the reference's FIR inner loops
// (filter.cpp:142-144,169-174,202-207).
__device__ __forceinline__ float xmac(float acc, float h, float x) {
  return __fadd_rn(acc, __fmul_rn(h, x));
}

// Two lanes of the same two-rounding multiply-add in one instruction pair (sm_100a FMUL2 + FFMA2):
// the product is rounded by mul.rn.f32x2, the sum by an fma whose multiplier is a RUN-TIME 1.0
// (p*1 is exact, so the fma rounds p + acc once).  ptxas contracts an explicit mul.rn.f32x2 /
// add.rn.f32x2 pair into one FFMA2 (seen with CUDA 12.9, -fmad=false or not), which would drop the
// product's rounding; it cannot contract through a multiplier it does not know.  Each lane is
// bit-identical to xmac; the pair takes two issue slots instead of four
// (tools/ubench_f32x2_issue.cu).
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float &lo, float &hi) {
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t xmac2(f32x2_t acc, f32x2_t h, f32x2_t x, f32x2_t one) {
  f32x2_t p;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(p) : "l"(h), "l"(x));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(acc) : "l"(p), "l"(one), "l"(acc));
  return acc;
}
// contracted pair (FAST / MIXED only)
__device__ __forceinline__ f32x2_t fma2(f32x2_t acc, f32x2_t h, f32x2_t x) {
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(acc) : "l"(h), "l"(x), "l"(acc));
  return acc;
}

// (u8 - 128) as an exact float without an integer->float conversion: the byte is
// placed in the low mantissa bits of 2^23 and the bias is subtracted.
__device__ __forceinline__ float u8_centered(uint32_t byte) {
  return __fsub_rn(__uint_as_float(0x4B000000u | byte), 8388736.0f);
}
// iofunc.cpp:128-135: float((u8-128)/128.0); exact, so the power-of-two scaling
// may be applied to the byte or (as the fused kernels do) to the filtered sum.
__device__ __forceinline__ float u8_to_unit(uint32_t byte) {
  return __fmul_rn(u8_centered(byte), 0.0078125f);
}

// threadMonoOnly.cpp:185-190: NaN -> 0, else static_cast<short>(v*16384).  On
// x86-64 the cast is cvttss2si (32-bit, "integer indefinite" 0x80000000 when out
// of range) followed by a 16-bit truncation; reproduce exactly that.
__device__ __forceinline__ int16_t pcm16(float v) {
  if (v != v) return 0;
  float t = __fmul_rn(v, 16384.0f);
  int32_t i = (t >= -2147483648.0f && t < 2147483648.0f) ? __float2int_rz(t) : INT32_MIN;
  return (int16_t)(uint16_t)((uint32_t)i & 0xffffu);
}

// fmDemod, filter.cpp:254-260, one sample.
__device__ __forceinline__ float fm_demod_one(float i, float q, float pi, float pq) {
  float den = xadd(xmul(i, i), xmul(q, q));
  if (den == 0.0f) return 0.0f;
  float num = xsub(xmul(i, xsub(q, pq)), xmul(q, xsub(i, pi)));
  return xdiv(num, den);
}

// Programmatic dependent launch (the mono fast path's three kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization): pdl_trigger lets the NEXT kernel of the stream be
// scheduled as soon as every CTA of this one has started -- its CTAs take the SM slots this kernel's
// CTAs free as they finish and run their prologue (barriers, tensor memory, tap tiles) under this
// kernel's tail --; pdl_wait blocks until the PREVIOUS kernel has completed and its writes are visible,
// and must precede the first access to anything that kernel produced.  Both are no-ops in a kernel
// launched the ordinary way.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- host-side error plumbing -------------------------------------------
void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

}  // namespace sdr

#define SDR_CUDA(call)                                                  \
  do {                                                                  \
    cudaError_t e__ = (call);                                           \
    if (e__ != cudaSuccess) return sdr::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)
