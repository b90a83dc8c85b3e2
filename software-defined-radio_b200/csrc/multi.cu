// multi.cu -- one process, N devices: the batched receiver spread over every B200 of the box.
//
// The reference runs one producer and one consumer thread around a bounded queue
// (src/project.cpp:471-496).  Here the unit of parallelism is the capture: `batch` independent
// captures are cut into contiguous ranges, one per device (captures never interact, so
// there is no collective), each range owned by a host worker thread that drives its own
// sdr_pipeline through sdr_pipeline_process_host (pinned double buffers, three CUDA streams).
// Every device writes its PCM straight into its rows of the caller's buffer, so the "final host
// gather" of the whole batch is the set of device-to-host copies themselves.
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sdr_b200.h"
#include "common.cuh"

namespace sdr {
int fail(int code, const std::string &msg);
}
using sdr::fail;

struct sdr_multi {
  struct Worker {
    int device = 0;
    int first = 0, count = 0;   // capture range [first, first + count)
    sdr_pipeline *pipe = nullptr;
    std::thread thread;
    // job slot (guarded by the owner's mutex)
    int job = 0;                // 0 idle, 1 process, 2 reset, 3 quit
    int rc = SDR_OK;
    std::string err;
    uint64_t launches = 0;
  };
  sdr_config cfg{};
  std::vector<Worker> workers;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  int pending = 0;
  // arguments of the job in flight
  const uint8_t *iq = nullptr;
  size_t iq_stride = 0, nbytes = 0;
  int16_t *pcm = nullptr;
  size_t pcm_stride = 0;
};

static void worker_main(sdr_multi *m, sdr_multi::Worker *w) {
  for (;;) {
    int job;
    {
      std::unique_lock<std::mutex> lk(m->mu);
      m->cv_job.wait(lk, [&] { return w->job != 0; });
      job = w->job;
    }
    int rc = SDR_OK;
    if (job == 1) {
      rc = sdr_pipeline_process_host(w->pipe, m->iq + (size_t)w->first * m->iq_stride, m->iq_stride, m->nbytes,
                                     m->pcm + (size_t)w->first * m->pcm_stride, m->pcm_stride);
    } else if (job == 2) {
      rc = sdr_pipeline_reset(w->pipe);
    }
    {
      std::unique_lock<std::mutex> lk(m->mu);
      w->rc = rc;
      if (rc) w->err = sdr_last_error();   // thread-local in the worker: hand it to the caller
      w->job = 0;
      if (--m->pending == 0) m->cv_done.notify_all();
    }
    if (job == 3) return;
  }
}

// Runs `job` on every worker and returns the first failure (message re-published on the calling thread).
static int run_all(sdr_multi *m, int job) {
  std::unique_lock<std::mutex> lk(m->mu);
  m->pending = (int)m->workers.size();
  for (auto &w : m->workers) w.job = job;
  m->cv_job.notify_all();
  m->cv_done.wait(lk, [&] { return m->pending == 0; });
  for (auto &w : m->workers)
    if (w.rc) return fail(w.rc, "device " + std::to_string(w.device) + ": " + w.err);
  return SDR_OK;
}

extern "C" int sdr_multi_create(const sdr_multi_config *mc, sdr_multi **out) {
  if (!mc || !out) return fail(SDR_ERR_INVALID, "null argument");
  *out = nullptr;
  const int usable = sdr_device_count();
  if (usable < 1) return fail(SDR_ERR_NO_DEVICE, "no sm_100 device available (this library has no CPU fallback)");
  int n = mc->n_devices > 0 ? mc->n_devices : usable;
  if (mc->cfg.batch < 1) return fail(SDR_ERR_INVALID, "batch must be >= 1");
  n = std::min(n, mc->cfg.batch);   // no device is left without a capture
  std::vector<int> devs(n);
  for (int i = 0; i < n; ++i) devs[i] = mc->devices ? mc->devices[i] : i;
  sdr_multi *m = new sdr_multi();
  m->cfg = mc->cfg;
  m->workers.resize(n);
  // contiguous ranges, sizes differing by at most one capture
  const int B = mc->cfg.batch;
  int first = 0;
  for (int i = 0; i < n; ++i) {
    auto &w = m->workers[i];
    w.device = devs[i];
    w.first = first;
    w.count = B / n + (i < B % n ? 1 : 0);
    first += w.count;
    sdr_config c = mc->cfg;
    c.batch = w.count;
    c.device = w.device;
    const int rc = sdr_pipeline_create(&c, &w.pipe);
    if (rc) {
      const std::string msg = sdr_last_error();
      for (auto &x : m->workers) sdr_pipeline_destroy(x.pipe);
      delete m;
      return fail(rc, "device " + std::to_string(devs[i]) + ": " + msg);
    }
  }
  for (auto &w : m->workers) w.thread = std::thread(worker_main, m, &w);
  *out = m;
  return SDR_OK;
}

extern "C" int sdr_multi_destroy(sdr_multi *m) {
  if (!m) return SDR_OK;
  run_all(m, 3);
  for (auto &w : m->workers) {
    if (w.thread.joinable()) w.thread.join();
    sdr_pipeline_destroy(w.pipe);
  }
  delete m;
  return SDR_OK;
}

extern "C" int sdr_multi_reset(sdr_multi *m) {
  if (!m) return fail(SDR_ERR_INVALID, "null handle");
  return run_all(m, 2);
}

extern "C" int sdr_multi_layout(const sdr_multi *m, int *n_devices, int *devices, int *first_capture, int cap) {
  if (!m || !n_devices) return fail(SDR_ERR_INVALID, "null argument");
  *n_devices = (int)m->workers.size();
  for (int i = 0; i < (int)m->workers.size() && i < cap; ++i) {
    if (devices) devices[i] = m->workers[i].device;
    if (first_capture) first_capture[i] = m->workers[i].first;
  }
  return SDR_OK;
}

extern "C" int sdr_multi_pcm_count(const sdr_multi *m, size_t nbytes, size_t *n_pcm) {
  if (!m) return fail(SDR_ERR_INVALID, "null handle");
  return sdr_pipeline_pcm_count(m->workers[0].pipe, nbytes, n_pcm);
}

extern "C" int sdr_multi_launch_count(sdr_multi *m, uint64_t *count, int reset) {
  if (!m) return fail(SDR_ERR_INVALID, "null handle");
  uint64_t total = 0;
  for (auto &w : m->workers) {
    uint64_t c = 0;
    const int rc = sdr_pipeline_launch_count(w.pipe, &c, reset);
    if (rc) return rc;
    total += c;
  }
  if (count) *count = total;
  return SDR_OK;
}

extern "C" int sdr_multi_process_host(sdr_multi *m, const uint8_t *iq, size_t iq_stride, size_t nbytes,
                                      int16_t *pcm, size_t pcm_stride) {
  if (!m || !iq || !pcm) return fail(SDR_ERR_INVALID, "null argument");
  if (iq_stride < nbytes) return fail(SDR_ERR_INVALID, "iq_stride smaller than nbytes_per_channel");
  m->iq = iq;
  m->iq_stride = iq_stride;
  m->nbytes = nbytes;
  m->pcm = pcm;
  m->pcm_stride = pcm_stride;
  return run_all(m, 1);
}
