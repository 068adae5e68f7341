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

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t smem_addr, uint32_t lbo) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

// flags: 1 = random operand bits instead of zeros (switching activity -> power), 2 = commit every group of four MMAs to a
// ring of four barriers and wait for the group issued four groups earlier, like the GEMM main loop does for its stages,
// 4 = rotate the operands through four 48 KB stage buffers instead of re-reading one, 8 = twelve more warps spin on an
// mbarrier for the whole loop (the GEMM's epilogue warps waiting for their accumulator), 16 = the stage hand-shake goes
// through a producer thread in another warp (empty -> producer -> full) as in the GEMM, 32 = warp-converged issue with
// elect.sync and 32-bit descriptor arithmetic (what csrc/gemm.cu does), commit ring included
__global__ void __launch_bounds__(448, 1) mma_bench_kernel(int N, int a_mn, int b_mn, int iters, long long *cycles, int flags) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t ring[4];
  __shared__ uint64_t fullb[4];
  __shared__ uint64_t pollb;
  __shared__ uint32_t slot;
  const int n_stage = (flags & 4) ? 4 : 1;
  for (int i = threadIdx.x; i < n_stage * (16384 + 32768) / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12;
    // bf16 pairs with exponents near 1.0 (no inf / nan): keep sign + mantissa bits random
    reinterpret_cast<uint32_t *>(smem)[i] = (flags & 1) ? ((h & 0x807F807Fu) | 0x3F003F00u) : 0u;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(ring + i, 1);
    for (int i = 0; i < 4; ++i) mbar_init(fullb + i, 1);
    mbar_init(&pollb, 1);
    mbar_fence_init();
  }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = slot;
  if ((flags & 32) && threadIdx.x < 32) {
    // the lean form: the whole warp runs the loop, elect.sync predicates the tcgen05 instructions (tc05.cuh)
    const uint32_t a_lo0 = desc_lo_sw128(smem_u32(smem), a_mn, 8192), b_lo0 = desc_lo_sw128(smem_u32(smem + 16384), b_mn, 8192);
    const uint32_t a_step = a_mn ? 2048 >> 4 : 32 >> 4, b_step = b_mn ? 2048 >> 4 : 32 >> 4;
    const uint32_t idesc = make_idesc_bf16(128, N) | (a_mn ? 1u << 15 : 0u) | (b_mn ? 1u << 16 : 0u);
    const uint32_t ring0 = smem_u32(ring);
    unsigned long long ns0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
    const long long t0 = clock64();
    uint32_t s = 0, ph = 1;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(ring + s, ph);   // first pass: the barrier's preceding phase reads as complete
      const uint32_t so = (flags & 4) ? s * (49152u >> 4) : 0u;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        mma_bf16_elect<1>(tmem, a_lo0 + so + kk * a_step, DESC_HI_SW128, b_lo0 + so + kk * b_step, DESC_HI_SW128, idesc, 1u);
      mma_commit_elect<1>(ring0 + s * 8);
      if (++s == 4) { s = 0; ph ^= 1u; }
    }
    mma_commit_elect<1>(smem_u32(&bar));
    mbar_wait(&bar, 0);
    if (threadIdx.x == 0) {
      cycles[blockIdx.x] = clock64() - t0;
      unsigned long long ns1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
      cycles[gridDim.x + blockIdx.x] = (long long)(ns1 - ns0);
      mbar_arrive(&pollb);
    }
  } else if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 16384);
    const uint32_t idesc = make_idesc_bf16(128, N) | (a_mn ? 1u << 15 : 0u) | (b_mn ? 1u << 16 : 0u);
    unsigned long long ns0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (flags & 16) {
        mbar_wait(fullb + (it & 3), (it >> 2) & 1u);
        fence_after_sync();
      } else if ((flags & 2) && it >= 4) {
        mbar_wait(ring + (it & 3), ((it >> 2) - 1) & 1u);
      }
      const uint32_t so = (flags & 4) ? (uint32_t)(it & 3) * 49152u : 0u;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t ad = a_mn ? desc_mn(a0 + so + kk * 2048, 8192) : make_desc_sw128(a0 + so + kk * 32);
        const uint64_t bd = b_mn ? desc_mn(b0 + so + kk * 2048, 8192) : make_desc_sw128(b0 + so + kk * 32);
        mma_bf16(tmem, ad, bd, idesc, 1u);
      }
      if (flags & (2 | 16)) mma_commit(ring + (it & 3));
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
    mbar_arrive(&pollb);
    unsigned long long ns1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
    cycles[gridDim.x + blockIdx.x] = (long long)(ns1 - ns0);
  }
  if (threadIdx.x == 32 && (flags & 16)) {   // producer: slot free (ring) -> operands "landed" (fullb)
    for (int it = 0; it < iters; ++it) {
      if (it >= 4) mbar_wait(ring + (it & 3), ((it >> 2) - 1) & 1u);
      mbar_arrive(fullb + (it & 3));
    }
  }
  if (threadIdx.x >= 64 && (flags & 8)) mbar_wait(&pollb, 0);
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}
}  // namespace

// cycles_out: [2][blocks] — SM cycles, then nanoseconds (globaltimer) of the issue loop of each block
extern "C" int sv_mma_bench(int N, int a_mn, int b_mn, int iters, int blocks, long long *cycles_out, void *stream) {
  const int flags = a_mn >> 1;   // bits 1.. of a_mn carry the profiling flags
  a_mn &= 1;
  if (N < 16 || N > 256 || (N % 16) || iters < 1 || blocks < 1 || !cycles_out) return SV_ERR_INVALID_ARG;
  static bool configured = false;
  if (!configured) {
    int rc = sv::cuda_status(cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 49152));
    if (rc) return rc;
    configured = true;
  }
  mma_bench_kernel<<<blocks, 448, (flags & 4) ? 4 * 49152 : 49152, (cudaStream_t)stream>>>(N, a_mn, b_mn, iters, cycles_out, flags);
  return sv::after_launch();
}
