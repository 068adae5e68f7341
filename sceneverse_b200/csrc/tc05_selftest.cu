// Single-tile tcgen05 GEMM used by tests/test_tc05_gpu.py to pin the descriptor / layout / TMEM
// conventions of tc05.cuh against torch.matmul:  D[128,N] (f32) = A[128,K] (bf16) x B[N,K]^T (bf16).
#include "svcommon.h"
#include "svgps.h"
#include "tc05.cuh"

namespace {

__global__ void __launch_bounds__(128) tc05_selftest_kernel(const __nv_bfloat16 *__restrict__ A,
                                                            const __nv_bfloat16 *__restrict__ B, float *__restrict__ D,
                                                            int N, int K, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t *sA = smem;
  uint8_t *sB = smem + 128 * K * 2;
  for (int e = tid; e < 128 * K; e += 128) {
    const int r = e / K, k = e - r * K;
    *reinterpret_cast<__nv_bfloat16 *>(sA + tc05::tile_off(128, r, k)) = A[e];
  }
  for (int e = tid; e < N * K; e += 128) {
    const int r = e / K, k = e - r * K;
    *reinterpret_cast<__nv_bfloat16 *>(sB + tc05::tile_off(N, r, k)) = B[e];
  }
  if (tid == 0) {
    tc05::mbar_init(&bar, 1);
    tc05::mbar_fence_init();
  }
  tc05::fence_proxy_async_smem();
  if (warp == 0) tc05::tmem_alloc<256>(&tmem_base_s);
  tc05::fence_before_sync();
  __syncthreads();
  tc05::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = tc05::make_idesc_bf16(128, N);
    const uint32_t a0 = tc05::smem_u32(sA), b0 = tc05::smem_u32(sB);
    const uint32_t kstrA = 128 * 16, kstrB = N * 16, rstr = 128;
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t ad = mode == 0 ? tc05::make_desc(a0 + ks * 2 * kstrA, kstrA, rstr)
                                    : tc05::make_desc(a0 + ks * 2 * kstrA, rstr, kstrA);
      const uint64_t bd = mode == 0 ? tc05::make_desc(b0 + ks * 2 * kstrB, kstrB, rstr)
                                    : tc05::make_desc(b0 + ks * 2 * kstrB, rstr, kstrB);
      tc05::mma_bf16(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
    }
    tc05::mma_commit(&bar);
  }
  tc05::mbar_wait(&bar, 0);
  tc05::fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tc05::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    float *drow = D + (size_t)(warp * 32 + lane) * N + c0;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (c0 + i < N) drow[i] = v[i];
  }
  tc05::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc05::tmem_dealloc<256>(tmem);
}

}  // namespace

extern "C" int sv_tc05_selftest(const void *A, const void *B, float *D, int N, int K, int mode, void *stream) {
  if (!A || !B || !D || N < 16 || N > 256 || (N % 16) || K < 16 || K > 256 || (K % 16)) return SV_ERR_INVALID_ARG;
  const size_t smem = (size_t)(128 + N) * K * 2;
  int rc = sv::cuda_status(
      cudaFuncSetAttribute(tc05_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (rc) return rc;
  tc05_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16 *)A, (const __nv_bfloat16 *)B, D,
                                                              N, K, mode);
  return sv::after_launch();
}
