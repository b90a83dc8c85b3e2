// h2d_ceiling.cu -- stand-alone copy-ceiling probe for the end-to-end leg (bench.py e2e).
// Plain page-locked host -> device copies with no pipeline behind them: one host thread and one
// stream per device, `slices` cudaMemcpyAsync of `slice_mb` MiB each per round, N = 1, 2, 4, 8
// devices concurrently (as many as the box has).  Prints one JSON line per N with the aggregate
// GB/s; what sdr_pipeline_process_host / sdr_multi_process_host reach is reported against it.
//   nvcc -O2 -o tools/h2d_ceiling.bin tools/h2d_ceiling.cu && tools/h2d_ceiling.bin [slice_mb] [slices]
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

int main(int argc, char **argv) {
  const size_t slice = (size_t)(argc > 1 ? atoi(argv[1]) : 64) << 20;
  const int slices = argc > 2 ? atoi(argv[2]) : 14;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    fprintf(stderr, "no CUDA device\n");
    return 1;
  }
  for (int n = 1; n <= n_dev; n *= 2) {
    std::vector<void *> h(n), d(n);
    std::vector<cudaStream_t> st(n);
    for (int i = 0; i < n; ++i) {
      cudaSetDevice(i);
      cudaMallocHost(&h[i], slice * slices);
      memset(h[i], 1, slice * slices);
      cudaMalloc(&d[i], slice * 2);
      cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
    }
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
      for (int i = 0; i < n; ++i) { cudaSetDevice(i); cudaDeviceSynchronize(); }
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> th;
      for (int i = 0; i < n; ++i)
        th.emplace_back([&, i] {
          cudaSetDevice(i);
          for (int k = 0; k < slices; ++k)
            cudaMemcpyAsync((char *)d[i] + (k & 1) * slice, (char *)h[i] + (size_t)k * slice, slice,
                            cudaMemcpyHostToDevice, st[i]);
          cudaStreamSynchronize(st[i]);
        });
      for (auto &t : th) t.join();
      const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      const double gbs = (double)n * slice * slices / sec / 1e9;
      if (rep && gbs > best) best = gbs;
    }
    printf("{\"devices\": %d, \"slice_mib\": %zu, \"slices\": %d, \"h2d_gbs_aggregate\": %.1f, \"h2d_gbs_per_device\": %.1f}\n",
           n, slice >> 20, slices, best, best / n);
    fflush(stdout);
    for (int i = 0; i < n; ++i) {
      cudaSetDevice(i);
      cudaFreeHost(h[i]);
      cudaFree(d[i]);
      cudaStreamDestroy(st[i]);
    }
  }
  return 0;
}
