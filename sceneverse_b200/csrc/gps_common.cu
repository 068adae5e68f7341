// libsvgps: process-wide bookkeeping shared by the tensor-core kernels (separate .so from libsvpointops).
#include <math.h>

#include "svcommon.h"

namespace sv {
std::atomic<unsigned long long> g_launches{0};
thread_local int t_last_cuda_error = 0;
const unsigned long long *g_seed_offset = nullptr;
int ref_opt_n_threads(int work_size) {
  const int pow_2 = (int)(std::log(static_cast<double>(work_size)) / std::log(2.0));
  int v = 1 << pow_2;
  if (v > 512) v = 512;
  return v < 1 ? 1 : v;
}
}  // namespace sv

extern "C" {
unsigned long long svgps_launch_count(void) { return sv::g_launches.load(); }
int svgps_last_cuda_error(void) { return sv::t_last_cuda_error; }
const char *svgps_last_cuda_error_string(void) { return cudaGetErrorString((cudaError_t)sv::t_last_cuda_error); }
int sv_dropout_seed_offset(const unsigned long long *device_counter) {
  sv::g_seed_offset = device_counter;
  return SV_OK;
}
}
