// pipeline_view.h -- what other translation units of the library (rds.cu) may see of a
// pipeline handle: its demodulated-IF buffer and a hook that runs at the end of every
// sdr_pipeline_process_device call, on that call's stream.
#pragma once

#include <cuda_runtime.h>

#include <string>

struct sdr_pipeline;

namespace sdr {

struct DemodView {
  const float *demod;  // [batch][stride]; this call's fm_demod starts at element `off` of a row
  size_t stride;
  int off;
  int batch, device, mode, if_Fs, rf_decim;
  size_t cap_if;  // IF samples per capture that one call may produce
};

// event 0: a process call produced n_if fm_demod samples per capture (kernels go on stream s);
// event 1: sdr_pipeline_reset;
// event 2: a process call of n_if samples is about to be enqueued (validate only);
// event 4: a host call of n_if samples in total is about to be cut into process calls (validate
//          what concerns the whole: room for its results).
typedef int (*PipelineHook)(void *ctx, int event, size_t n_if, cudaStream_t s);

int pipeline_view(sdr_pipeline *p, DemodView *v);
// granule_bytes: the pipeline's granule becomes the least common multiple of its own and this.
int pipeline_set_hook(sdr_pipeline *p, PipelineHook fn, void *ctx, int granule_bytes);
int fail(int code, const std::string &msg);

}  // namespace sdr

// launch bookkeeping shared with the pipeline (gpu_launches, per-kernel event timing)
void sdr_prof_begin(sdr_pipeline *p, const char *name, cudaStream_t s);
int sdr_check_launch(sdr_pipeline *p, const char *name);
