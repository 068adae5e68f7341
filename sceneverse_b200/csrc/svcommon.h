// Shared host-side plumbing of libsvpointops / libsvgps (status codes, launch bookkeeping).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "svpointops.h"

namespace sv {

extern std::atomic<unsigned long long> g_launches;
extern thread_local int t_last_cuda_error;
// device counter added (on the device, at run time) to the seed of every in-kernel dropout mask; see sv_dropout_seed_offset
extern const unsigned long long *g_seed_offset;

// Call after every <<<>>>: records the launch, maps a launch failure to SV_ERR_CUDA.
inline int after_launch() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    t_last_cuda_error = (int)e;
    return SV_ERR_CUDA;
  }
  return SV_OK;
}

inline int cuda_status(cudaError_t e) {
  if (e != cudaSuccess) {
    t_last_cuda_error = (int)e;
    return SV_ERR_CUDA;
  }
  return SV_OK;
}

// include/cuda_utils.h:13-19 of the reference: the block size the reference launches with decides
// the FPS tie-break order, so it is reproduced with the same libm expression.
int ref_opt_n_threads(int work_size);

// csrc/fps_coop.cu: multi-CTA register-resident FPS for 8192 < N; SV_ERR_INVALID_ARG = shape not supported by this path
int fps_coop(const float *xyz, int B, int N, int m, int *idx, float *new_xyz, cudaStream_t st);

}  // namespace sv
