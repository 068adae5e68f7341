/*
 * libsvgps — C-ABI of the tensor-core (tcgen05) kernels of the GPS object encoder and attention stack.
 * Raw device pointers + sizes + cudaStream_t (void*), int status (codes in svpointops.h), no
 * allocation, no synchronisation.  bf16 tensors are passed as void* (device, 2-byte elements).
 */
#ifndef SVGPS_H
#define SVGPS_H
#include "svpointops.h"

#ifdef __cplusplus
extern "C" {
#endif

/* bookkeeping of this library (separate counters from libsvpointops) */
unsigned long long svgps_launch_count(void);
int svgps_last_cuda_error(void);
const char *svgps_last_cuda_error_string(void);

/* Conformance probe of the tcgen05 conventions (tests only): D[128,N] f32 = A[128,K] bf16 x B[N,K]^T bf16,
 * one CTA, one TMEM accumulator; mode 0 = the descriptor convention used by the library. */
int sv_tc05_selftest(const void *A, const void *B, float *D, int N, int K, int mode, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SVGPS_H */
