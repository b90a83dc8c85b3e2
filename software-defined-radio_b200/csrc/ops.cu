// ops.cu -- the reference's filter.h functions one by one on the device, for
// callers that re-point individual calls (include/dropin/filter.h) and for the
// per-operator parity tests.  Each call uploads its operands, runs the same
// generic kernels the pipeline uses for unusual tap counts, and downloads the
// result; throughput work belongs in sdr_pipeline_*.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "../../include/sdr_b200.h"
#include "kernels.cuh"

using namespace sdr;

namespace {

int bad(const char *msg) {
  set_error(msg);
  return SDR_ERR_INVALID;
}

int select_device(int dev) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("no CUDA device available (this library has no CPU fallback)");
    return SDR_ERR_NO_DEVICE;
  }
  if (dev < 0 || dev >= n) return bad("device ordinal out of range");
  cudaDeviceProp p;
  SDR_CUDA(cudaGetDeviceProperties(&p, dev));
  if (p.major != 10) {
    set_error("device is not sm_100 (Blackwell B200); kernels are built for sm_100a only");
    return SDR_ERR_NO_DEVICE;
  }
  SDR_CUDA(cudaSetDevice(dev));
  return SDR_OK;
}

struct Tmp {
  void *p = nullptr;
  ~Tmp() {
    if (p) cudaFree(p);
  }
  template <typename T>
  T *as() { return static_cast<T *>(p); }
};

int dalloc(Tmp &t, size_t bytes) {
  SDR_CUDA(cudaMalloc(&t.p, std::max<size_t>(bytes, 16)));
  return SDR_OK;
}

int launch_ok(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, what, __FILE__, __LINE__);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return cuda_fail(e, what, __FILE__, __LINE__);
  return SDR_OK;
}

// y[j] = sum_n h[n] * X[j*D - n] with X = state ++ x; then state <- tail.
int fir_common(int device, float *y, const float *x, size_t nx, const float *h, size_t nh,
               float *state, unsigned decim) {
  if (!y || !x || !h || !state) return bad("null argument");
  if (nh < 1 || decim < 1) return bad("need at least one tap and decim >= 1");
  int rc = select_device(device);
  if (rc) return rc;
  const size_t ns = nh - 1, ny = nx / decim;
  Tmp dx, dh, dy;
  if ((rc = dalloc(dx, (ns + nx) * sizeof(float)))) return rc;
  if ((rc = dalloc(dh, nh * sizeof(float)))) return rc;
  if ((rc = dalloc(dy, ny * sizeof(float)))) return rc;
  SDR_CUDA(cudaMemcpy(dx.as<float>(), state, ns * sizeof(float), cudaMemcpyHostToDevice));
  SDR_CUDA(cudaMemcpy(dx.as<float>() + ns, x, nx * sizeof(float), cudaMemcpyHostToDevice));
  SDR_CUDA(cudaMemcpy(dh.p, h, nh * sizeof(float), cudaMemcpyHostToDevice));
  if (ny) {
    FirGenericArgs g{dx.as<float>(), 0, (int)ns, dh.as<float>(), (int)nh, (int)decim,
                     dy.as<float>(), 0, 0, (int)ny};
    k_fir_generic<<<dim3((unsigned)((ny + 127) / 128), 1), 128, nh * sizeof(float)>>>(g);
    if ((rc = launch_ok("k_fir_generic"))) return rc;
    SDR_CUDA(cudaMemcpy(y, dy.p, ny * sizeof(float), cudaMemcpyDeviceToHost));
  }
  // filter.cpp:148-153 / :183-187: the last nh-1 inputs become the next state
  // (taken from state ++ x so that blocks shorter than the filter also work).
  std::vector<float> cat(ns + nx);
  std::memcpy(cat.data(), state, ns * sizeof(float));
  std::memcpy(cat.data() + ns, x, nx * sizeof(float));
  std::memcpy(state, cat.data() + nx, ns * sizeof(float));
  return SDR_OK;
}

__global__ void k_convolve(const float *x, int nx, const float *h, int nh, float *y, int ny) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= ny) return;
  float acc = 0.0f;
  for (int n = 0; n < nh; ++n) {
    const int i = m - n;
    if (i >= 0 && i < nx) acc = xmac(acc, h[n], x[i]);  // filter.cpp:125-127
  }
  y[m] = acc;
}

__global__ void k_upsample(const float *x, int n_out, int up, float *xu) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_out) xu[i] = (i % up == 0) ? x[i / up] : 0.0f;  // filter.cpp:230-233
}

__global__ void k_downsample(const float *in, int n_out, int ds, float *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_out) out[i] = in[(size_t)i * ds];  // filter.cpp:241-244
}

__global__ void k_allpass(const float *in, int n, const float *state, int ns, float *out,
                          float *state_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (i < ns) ? state[i] : in[i - ns];      // filter.cpp:21,27-28
  if (i < ns) state_out[i] = in[n - ns + i];                  // filter.cpp:23-25
}

}  // namespace

extern "C" int sdr_convolve(int device, float *y, const float *x, size_t nx, const float *h,
                            size_t nh) {
  if (!y || !x || !h || !nx || !nh) return bad("null or empty argument");
  int rc = select_device(device);
  if (rc) return rc;
  const size_t ny = nx + nh - 1;
  Tmp dx, dh, dy;
  if ((rc = dalloc(dx, nx * 4)) || (rc = dalloc(dh, nh * 4)) || (rc = dalloc(dy, ny * 4))) return rc;
  SDR_CUDA(cudaMemcpy(dx.p, x, nx * 4, cudaMemcpyHostToDevice));
  SDR_CUDA(cudaMemcpy(dh.p, h, nh * 4, cudaMemcpyHostToDevice));
  k_convolve<<<(unsigned)((ny + 127) / 128), 128>>>(dx.as<float>(), (int)nx, dh.as<float>(), (int)nh,
                                                    dy.as<float>(), (int)ny);
  if ((rc = launch_ok("k_convolve"))) return rc;
  SDR_CUDA(cudaMemcpy(y, dy.p, ny * 4, cudaMemcpyDeviceToHost));
  return SDR_OK;
}

extern "C" int sdr_fir_block(int device, float *y, const float *x, size_t nx, const float *h,
                             size_t nh, float *state) {
  return fir_common(device, y, x, nx, h, nh, state, 1);
}

extern "C" int sdr_fir_decim(int device, float *y, const float *x, size_t nx, const float *h,
                             size_t nh, float *state, unsigned decim) {
  return fir_common(device, y, x, nx, h, nh, state, decim);
}

extern "C" int sdr_fir_resample(int device, float *y, const float *x, size_t nx, const float *h,
                                size_t nh, float *state, unsigned decim, unsigned upsamp) {
  if (!y || !x || !h || !state) return bad("null argument");
  if (!decim || !upsamp || nh < upsamp) return bad("need decim, upsamp >= 1 and at least U taps");
  if (nh % upsamp) return bad("tap count must be a multiple of upsamp (the reference uses taps*U)");
  const int U = (int)upsamp, TA = (int)(nh / upsamp);
  if (nx < (size_t)TA) return bad("block shorter than the per-phase filter");
  int rc = select_device(device);
  if (rc) return rc;
  const size_t ny = nx * upsamp / decim;
  // phase-major taps and the compact history hidden in the zero-stuffed state:
  // slot (j+1)*U-1 of `state` is input sample j-(TA-1) relative to this block.
  std::vector<float> hp((size_t)U * TA), xin((size_t)(TA - 1) + nx);
  for (int p = 0; p < U; ++p)
    for (int k = 0; k < TA; ++k) hp[(size_t)p * TA + k] = h[p + (size_t)k * U];
  for (int j = 0; j < TA - 1; ++j) xin[j] = state[(size_t)(j + 1) * U - 1];
  std::memcpy(xin.data() + (TA - 1), x, nx * sizeof(float));
  Tmp dx, dh, dy;
  if ((rc = dalloc(dx, xin.size() * 4)) || (rc = dalloc(dh, hp.size() * 4)) || (rc = dalloc(dy, ny * 4)))
    return rc;
  SDR_CUDA(cudaMemcpy(dx.p, xin.data(), xin.size() * 4, cudaMemcpyHostToDevice));
  SDR_CUDA(cudaMemcpy(dh.p, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice));
  if (ny) {
    ResampleOpArgs g{dx.as<float>(), TA - 1, dh.as<float>(), U, (int)decim, TA, dy.as<float>(), (int)ny};
    k_resample_op<<<(unsigned)((ny + 127) / 128), 128>>>(g);
    if ((rc = launch_ok("k_resample_op"))) return rc;
    SDR_CUDA(cudaMemcpy(y, dy.p, ny * 4, cudaMemcpyDeviceToHost));
  }
  // filter.cpp:217-222: only every U-th slot is rewritten; the rest stay as they were.
  for (int j = 0; j < TA - 1; ++j) state[(size_t)(j + 1) * U - 1] = x[nx - (TA - 1) + j];
  return SDR_OK;
}

extern "C" int sdr_fm_demod(int device, float *out, const float *I, const float *Q, size_t n,
                            float *prev_i, float *prev_q) {
  if (!out || !I || !Q || !prev_i || !prev_q) return bad("null argument");
  if (!n) return SDR_OK;
  int rc = select_device(device);
  if (rc) return rc;
  Tmp diq, dout;
  if ((rc = dalloc(diq, 2 * (n + 1) * 4)) || (rc = dalloc(dout, n * 4))) return rc;
  float *d = diq.as<float>();
  SDR_CUDA(cudaMemcpy(d, prev_i, 4, cudaMemcpyHostToDevice));
  SDR_CUDA(cudaMemcpy(d + 1, I, n * 4, cudaMemcpyHostToDevice));
  SDR_CUDA(cudaMemcpy(d + n + 1, prev_q, 4, cudaMemcpyHostToDevice));
  SDR_CUDA(cudaMemcpy(d + n + 2, Q, n * 4, cudaMemcpyHostToDevice));
  k_fm_demod<<<dim3((unsigned)((n + 127) / 128), 1), 128>>>(d, (int)n, dout.as<float>(), 0, 0);
  if ((rc = launch_ok("k_fm_demod"))) return rc;
  SDR_CUDA(cudaMemcpy(out, dout.p, n * 4, cudaMemcpyDeviceToHost));
  *prev_i = I[n - 1];  // filter.cpp:264-265
  *prev_q = Q[n - 1];
  return SDR_OK;
}

extern "C" int sdr_pll(int device, const float *in, size_t n, float *out, float *state, float freq,
                       float Fs, float ncoScale, float phaseAdjust, float normBandwidth) {
  if (!in || !out || !state) return bad("null argument");
  int rc = select_device(device);
  if (rc) return rc;
  if (n == 0) {  // empty PLLIn: ncoOut = {state[4]}, state unchanged (filter.cpp:41-46,73-79)
    out[0] = state[4];
    return SDR_OK;
  }
  Tmp din, dout, dst;
  if ((rc = dalloc(din, n * 4)) || (rc = dalloc(dout, (n + 1) * 4)) || (rc = dalloc(dst, 8 * 4))) return rc;
  float st8[8] = {state[0], state[1], state[2], state[3], state[4], state[5], 0, 0};
  SDR_CUDA(cudaMemcpy(din.p, in, n * 4, cudaMemcpyHostToDevice));
  SDR_CUDA(cudaMemcpy(dst.p, st8, sizeof st8, cudaMemcpyHostToDevice));
  SDR_CUDA(cudaMemcpy(dout.p, &state[4], 4, cudaMemcpyHostToDevice));  // ncoOut[0] = state[4]
  PllArgs a{din.as<float>(), 0, dout.as<float>(), 0, 0, dst.as<float>(), (int)n, 1,
            freq, Fs, ncoScale, phaseAdjust, normBandwidth};
  k_pll<<<1, 32>>>(a);
  if ((rc = launch_ok("k_pll"))) return rc;
  k_nco_cos<<<dim3(((unsigned)n + 1023) / 1024, 1), 256>>>(a);
  if ((rc = launch_ok("k_nco_cos"))) return rc;
  SDR_CUDA(cudaMemcpy(out, dout.p, (n + 1) * 4, cudaMemcpyDeviceToHost));
  SDR_CUDA(cudaMemcpy(st8, dst.p, sizeof st8, cudaMemcpyDeviceToHost));
  std::memcpy(state, st8, 6 * sizeof(float));
  return SDR_OK;
}

extern "C" int sdr_allpass(int device, const float *in, size_t n, float *state, size_t ns,
                           float *out) {
  if (!in || !state || !out) return bad("null argument");
  if (ns > n) return bad("all-pass state longer than the block");
  int rc = select_device(device);
  if (rc) return rc;
  Tmp din, dst, dout, dst2;
  if ((rc = dalloc(din, n * 4)) || (rc = dalloc(dst, ns * 4)) || (rc = dalloc(dout, n * 4)) ||
      (rc = dalloc(dst2, ns * 4)))
    return rc;
  SDR_CUDA(cudaMemcpy(din.p, in, n * 4, cudaMemcpyHostToDevice));
  SDR_CUDA(cudaMemcpy(dst.p, state, ns * 4, cudaMemcpyHostToDevice));
  k_allpass<<<(unsigned)((n + 127) / 128), 128>>>(din.as<float>(), (int)n, dst.as<float>(), (int)ns,
                                                  dout.as<float>(), dst2.as<float>());
  if ((rc = launch_ok("k_allpass"))) return rc;
  SDR_CUDA(cudaMemcpy(out, dout.p, n * 4, cudaMemcpyDeviceToHost));
  SDR_CUDA(cudaMemcpy(state, dst2.p, ns * 4, cudaMemcpyDeviceToHost));
  return SDR_OK;
}

extern "C" int sdr_upsample(int device, const float *x, size_t nx, float *xu, int up_rate) {
  if (!x || !xu || up_rate < 1) return bad("bad argument");
  int rc = select_device(device);
  if (rc) return rc;
  const size_t ny = nx * (size_t)up_rate;
  Tmp dx, dy;
  if ((rc = dalloc(dx, nx * 4)) || (rc = dalloc(dy, ny * 4))) return rc;
  SDR_CUDA(cudaMemcpy(dx.p, x, nx * 4, cudaMemcpyHostToDevice));
  k_upsample<<<(unsigned)((ny + 127) / 128), 128>>>(dx.as<float>(), (int)ny, up_rate, dy.as<float>());
  if ((rc = launch_ok("k_upsample"))) return rc;
  SDR_CUDA(cudaMemcpy(xu, dy.p, ny * 4, cudaMemcpyDeviceToHost));
  return SDR_OK;
}

extern "C" int sdr_downsample(int device, float *out, const float *in, size_t n, unsigned short ds) {
  if (!out || !in || ds < 1) return bad("bad argument");
  int rc = select_device(device);
  if (rc) return rc;
  const size_t ny = (n + ds - 1) / ds;  // filter.cpp:240: ceil(n / ds)
  Tmp dx, dy;
  if ((rc = dalloc(dx, n * 4)) || (rc = dalloc(dy, ny * 4))) return rc;
  SDR_CUDA(cudaMemcpy(dx.p, in, n * 4, cudaMemcpyHostToDevice));
  k_downsample<<<(unsigned)((ny + 127) / 128), 128>>>(dx.as<float>(), (int)ny, ds, dy.as<float>());
  if ((rc = launch_ok("k_downsample"))) return rc;
  SDR_CUDA(cudaMemcpy(out, dy.p, ny * 4, cudaMemcpyDeviceToHost));
  return SDR_OK;
}
