// Micro-benchmark (tests / profiling only): issue rate of tcgen05.mma by operand layout.  One CTA per SM, operands resident
// in shared memory (zero-filled), one thread issues `iters` x 4 MMAs (K = 64 per group) and waits for the commit; reports
// SM cycles per MMA.  Answers one question for csrc/gemm.cu: what does an MN-major (transposed) shared-memory operand cost
// relative to a K-major one at a given tile width?
#include <cuda.h>
#include <cuda_bf16.h>

#include "svcommon.h"
#include "svgps.h"
#include "tc05.cuh"

namespace {
using namespace tc05;

__device__ __forceinline__ uint64_t desc_mn(uint32_t smem_addr, uint32_t lbo) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int N, int a_mn, int b_mn, int iters, long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 16384);
    const uint32_t idesc = make_idesc_bf16(128, N) | (a_mn ? 1u << 15 : 0u) | (b_mn ? 1u << 16 : 0u);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t ad = a_mn ? desc_mn(a0 + kk * 2048, 8192) : make_desc_sw128(a0 + kk * 32);
        const uint64_t bd = b_mn ? desc_mn(b0 + kk * 2048, 8192) : make_desc_sw128(b0 + kk * 32);
        mma_bf16(tmem, ad, bd, idesc, 1u);
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}
}  // namespace

extern "C" int sv_mma_bench(int N, int a_mn, int b_mn, int iters, int blocks, long long *cycles_out, void *stream) {
  if (N < 16 || N > 256 || (N % 16) || iters < 1 || blocks < 1 || !cycles_out) return SV_ERR_INVALID_ARG;
  static bool configured = false;
  if (!configured) {
    int rc = sv::cuda_status(cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152));
    if (rc) return rc;
    configured = true;
  }
  mma_bench_kernel<<<blocks, 128, 49152, (cudaStream_t)stream>>>(N, a_mn, b_mn, iters, cycles_out);
  return sv::after_launch();
}
