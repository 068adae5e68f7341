// Persistent, warp-specialised tcgen05 GEMM with fused epilogues — the dense-contraction workhorse of the GPS path
// (every nn.Linear of the attention stack / heads and the SA3 + fc stage of PointNet++; reference: F.linear /
// 1x1 Conv2d call sites listed in SURVEY.md §2.3).
//
//   C[M,N] = epilogue( A[M,K] (bf16, row-major)  x  B[N,K]^T (bf16, row-major = nn.Linear weight layout) )
//
// Roles (320 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..9 = epilogue (two warps per TMEM lane quadrant, each thread one accumulator row x half of the columns).  Operand tiles travel global -> shared by TMA
// (cp.async.bulk.tensor.2d, one box of [rows x 64 elements] = 128-byte rows, SWIZZLE_128B, read back by the tensor core
// through a SWIZZLE_128B K-major UMMA descriptor) through a 4-stage mbarrier ring; accumulators are double-buffered
// in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.  Out-of-bounds rows / columns / K are zero-filled by
// TMA, the epilogue masks its stores.
#include <cuda.h>
#include <cuda_bf16.h>

#include "svcommon.h"
#include "svgps.h"
#include "tc05.cuh"

namespace {

using namespace tc05;

constexpr int BM = 128, BK = 64, STAGES = 4;
constexpr int STG_BYTES = 32 * 80;  // per epilogue warp: 32 rows x (64 B + 16 B pad); also holds the [32][17] f32 row-max scratch

struct GemmArgs {
  int M, N, K;
  const float *bias;        // [N] or null
  const void *residual;     // [M,N] same dtype as out, or null (added after the activation)
  void *out;                // [M,N] (or [M/rowmax,N] when rowmax > 0)
  int act;                  // 0 none, 1 relu, 2 gelu(erf)
  int out_f32;              // 0 bf16, 1 f32
  int rowmax;               // 0, or 16: max over groups of 16 consecutive rows (SA3 neighbourhood max)
  int ldo;                  // leading dimension of out / residual (elements)
  int a_mn, b_mn;           // 1: the operand is stored [K][rows] (row-major, rows contiguous) = MN-major for the tensor core
};

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
// MN-major SWIZZLE_128B operand staged as 64-column slabs [BK rows][128 B] (one TMA box each): 8-row groups 1024 B apart
// (stride byte offset), consecutive 64-element MN blocks one slab (BK * 128 B) apart (leading byte offset); a K step of 16
// rows advances the start address by 2048 B.  Same descriptor family the attention kernels use for P.V / dS.K.
__device__ __forceinline__ uint64_t make_desc_sw128_mn(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((BK * 128) >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

template <int BN>
__global__ void __launch_bounds__(320, 1)
gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const GemmArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int A_STAGE = BM * BK * 2, B_STAGE = BN * BK * 2;
  uint8_t *sA = smem;
  uint8_t *sB = smem + STAGES * A_STAGE;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * (A_STAGE + B_STAGE));
  uint64_t *empty = full + STAGES;
  uint64_t *acc_full = empty + STAGES;   // [2]
  uint64_t *acc_empty = acc_full + 2;    // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
  float *stage = reinterpret_cast<float *>(tmem_slot + 4);  // [8 warps][STG_BYTES] store staging / row-max transpose scratch
  float *sbias = stage + 8 * STG_BYTES / 4;                     // [2][BN] bias slice of the tile, double-buffered

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (g.M + BM - 1) / BM, tiles_n = (g.N + BN - 1) / BN;
  const int n_tiles = tiles_m * tiles_n;
  const int k_steps = (g.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full + b, 1);
      mbar_init(acc_empty + b, 256);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<2 * BN>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int m0 = (t % tiles_m) * BM, n0 = (t / tiles_m) * BN;
        for (int ks = 0; ks < k_steps; ++ks, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(empty + s, ph ^ 1u);  // slot free (first pass returns immediately)
          mbar_expect_tx(full + s, A_STAGE + B_STAGE);
          const int k0 = ks * BK;
          if (g.a_mn) {  // [K][M] storage: one [BK x 64] box per 64 rows of the tile
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sA + s * A_STAGE + j * (BK * 128), &mapA, m0 + j * 64, k0, full + s);
          } else {
            tma_load_2d(sA + s * A_STAGE, &mapA, k0, m0, full + s);
          }
          if (g.b_mn) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sB + s * B_STAGE + j * (BK * 128), &mapB, n0 + j * 64, k0, full + s);
          } else {
            tma_load_2d(sB + s * B_STAGE, &mapB, k0, n0, full + s);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    if (lane == 0) {
      const uint32_t IDESC = make_idesc_bf16(BM, BN) | (g.a_mn ? 1u << 15 : 0u) | (g.b_mn ? 1u << 16 : 0u);
      uint32_t it = 0, tl = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tl) {
        const int b = tl & 1;
        mbar_wait(acc_empty + b, ((tl >> 1) & 1u) ^ 1u);  // epilogue drained this accumulator
        fence_after_sync();
        for (int ks = 0; ks < k_steps; ++ks, ++it) {
          const int s = it % STAGES;
          mbar_wait(full + s, (it / STAGES) & 1u);
          fence_after_sync();
          const uint32_t a0 = smem_u32(sA + s * A_STAGE), b0 = smem_u32(sB + s * B_STAGE);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk)
            mma_bf16(tmem + b * BN, g.a_mn ? make_desc_sw128_mn(a0 + kk * 2048) : make_desc_sw128(a0 + kk * 32),
                     g.b_mn ? make_desc_sw128_mn(b0 + kk * 2048) : make_desc_sw128(b0 + kk * 32), IDESC, (ks | kk) != 0);
          mma_commit(empty + s);  // frees the smem slot when these MMAs retire
        }
        mma_commit(acc_full + b);
      }
    }
  } else {
    // ------------------------------- epilogue (warps 2..9) -------------------------------
    // Two warps per TMEM lane quadrant, each taking half of the tile's columns.  Per tile the bias slice is staged once in
    // shared memory (double-buffered with the accumulator), TMEM loads run one 32-column chunk ahead of the math.
    const int q = warp & 3;             // TMEM lane quadrant this warp may access (hardware rule: warp id % 4)
    const int half = (warp - 2) >> 2;   // which half of the BN columns
    const int et = threadIdx.x - 64;    // 0..255
    const int row_in_tile = q * 32 + lane;
    constexpr int NCHUNK = BN / 64;     // 32-column chunks per warp
    uint32_t tl = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tl) {
      const int b = tl & 1;
      const int m0 = (t % tiles_m) * BM, n0 = (t / tiles_m) * BN;
      const int row = m0 + row_in_tile;
      float *sb = sbias + b * BN;
      if (et < BN) sb[et] = (g.bias != nullptr && n0 + et < g.N) ? __ldg(g.bias + n0 + et) : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");  // bias visible; everybody is done with the tile before last
      mbar_wait(acc_full + b, (tl >> 1) & 1u);
      fence_after_sync();
      const uint32_t taddr = tmem + b * BN + half * (BN / 2) + ((uint32_t)(q * 32) << 16);
      uint32_t rr[2][32];
      tmem_ld32_async(taddr, rr[0]);
#pragma unroll
      for (int ci = 0; ci < NCHUNK; ++ci) {
        uint32_t(&cur)[32] = rr[ci & 1];
        tmem_wait32(cur);
        if (ci + 1 < NCHUNK) tmem_ld32_async(taddr + (ci + 1) * 32, rr[(ci + 1) & 1]);
        const int c0 = half * (BN / 2) + ci * 32;
        const int col0 = n0 + c0;
        if (col0 < g.N) {
        float v[32];
        const float4 *b4 = reinterpret_cast<const float4 *>(sb + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bb = b4[i];
          v[4 * i] = __uint_as_float(cur[4 * i]) + bb.x;
          v[4 * i + 1] = __uint_as_float(cur[4 * i + 1]) + bb.y;
          v[4 * i + 2] = __uint_as_float(cur[4 * i + 2]) + bb.z;
          v[4 * i + 3] = __uint_as_float(cur[4 * i + 3]) + bb.w;
        }
        if (g.act == 1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        } else if (g.act == 2) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
        }
        if (g.rowmax == 16) {
          // max over the 16 rows of each half-warp (16 consecutive rows = the 16 points of one cloud): 16 columns at a
          // time go through a per-warp [32][17] transpose scratch; lanes 0-15 reduce rows 0-15, lanes 16-31 rows 16-31
          float *tp = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(stage) + (warp - 2) * STG_BYTES);
          const int l16 = lane & 15, grp = lane >> 4;
          const int wrow0 = m0 + q * 32;  // first accumulator row of this warp
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
            for (int i = 0; i < 16; ++i) tp[lane * 17 + i] = row < g.M ? v[hh * 16 + i] : -INFINITY;
            __syncwarp();
            float mx = -INFINITY;
#pragma unroll
            for (int r2 = 0; r2 < 16; ++r2) mx = fmaxf(mx, tp[(grp * 16 + r2) * 17 + l16]);
            __syncwarp();
            const int col = col0 + hh * 16 + l16;
            if (col < g.N && wrow0 + grp * 16 < g.M) {
              const size_t orow = (size_t)(wrow0 >> 4) + grp;
              if (g.out_f32) reinterpret_cast<float *>(g.out)[orow * g.ldo + col] = mx;
              else reinterpret_cast<__nv_bfloat16 *>(g.out)[orow * g.ldo + col] = __float2bfloat16_rn(mx);
            }
          }
        } else {
          // Coalesced stores: the warp's 32 x 32 chunk goes through a per-warp [32][80 B] staging buffer (one row per
          // lane in, 8 rows x 64 B per store instruction out), so a store instruction touches 8 lines instead of 32.
          const bool full32 = col0 + 32 <= g.N;
          uint8_t *stg = reinterpret_cast<uint8_t *>(stage) + (warp - 2) * STG_BYTES;
          const int wrow0 = m0 + q * 32;
          const int sr = lane >> 2, seg = lane & 3;
          const bool vec_ok = full32 && g.residual == nullptr && (reinterpret_cast<uintptr_t>(g.out) & 15) == 0;
          if (g.out_f32) {
            float *ob = reinterpret_cast<float *>(g.out);
            if (vec_ok && (g.ldo & 3) == 0 && (col0 & 3) == 0) {
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  *reinterpret_cast<float4 *>(stg + lane * 80 + k * 16) =
                      make_float4(v[hh * 16 + 4 * k], v[hh * 16 + 4 * k + 1], v[hh * 16 + 4 * k + 2], v[hh * 16 + 4 * k + 3]);
                __syncwarp();
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                  const int r = it * 8 + sr;
                  const float4 w = *reinterpret_cast<const float4 *>(stg + r * 80 + seg * 16);
                  if (wrow0 + r < g.M)
                    *reinterpret_cast<float4 *>(ob + (size_t)(wrow0 + r) * g.ldo + col0 + hh * 16 + seg * 4) = w;
                }
                __syncwarp();
              }
            } else if (row < g.M) {
              float *o = ob + (size_t)row * g.ldo + col0;
              const float *r = g.residual ? reinterpret_cast<const float *>(g.residual) + (size_t)row * g.ldo + col0 : nullptr;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < g.N) o[i] = v[i] + (r ? r[i] : 0.f);
            }
          } else {
            __nv_bfloat16 *ob = reinterpret_cast<__nv_bfloat16 *>(g.out);
            if (vec_ok && (g.ldo & 7) == 0 && (col0 & 7) == 0) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                *reinterpret_cast<uint4 *>(stg + lane * 80 + k * 16) =
                    make_uint4(pack_bf16(v[8 * k], v[8 * k + 1]), pack_bf16(v[8 * k + 2], v[8 * k + 3]),
                               pack_bf16(v[8 * k + 4], v[8 * k + 5]), pack_bf16(v[8 * k + 6], v[8 * k + 7]));
              __syncwarp();
#pragma unroll
              for (int it = 0; it < 4; ++it) {
                const int r = it * 8 + sr;
                const uint4 w = *reinterpret_cast<const uint4 *>(stg + r * 80 + seg * 16);
                if (wrow0 + r < g.M) *reinterpret_cast<uint4 *>(ob + (size_t)(wrow0 + r) * g.ldo + col0 + seg * 8) = w;
              }
              __syncwarp();
            } else if (row < g.M) {
              __nv_bfloat16 *o = ob + (size_t)row * g.ldo + col0;
              const __nv_bfloat16 *r =
                  g.residual ? reinterpret_cast<const __nv_bfloat16 *>(g.residual) + (size_t)row * g.ldo + col0 : nullptr;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < g.N) o[i] = __float2bfloat16_rn(v[i] + (r ? __bfloat162float(r[i]) : 0.f));
            }
          }
        }
        }
      }
      fence_before_sync();
      mbar_arrive(acc_empty + b);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<2 * BN>(tmem);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// row-major [rows, K] bf16 with leading dimension ld (elements); box = [box_rows x 64 elements], 128-byte swizzle
int make_map(CUtensorMap *map, const void *ptr, int rows, int K, int ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return SV_ERR_CUDA;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};  // 64 bf16 = one 128-byte swizzle row
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SV_OK : SV_ERR_INVALID_ARG;
}

// [K, rows] bf16 row-major with leading dimension ld (the operand's rows are the contiguous direction): box = [BK x 64]
int make_map_mn(CUtensorMap *map, const void *ptr, int rows, int K, int ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return SV_ERR_CUDA;
  cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)K};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)BK};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SV_OK : SV_ERR_INVALID_ARG;
}

template <int BN>
int launch_gemm(const CUtensorMap &ma, const CUtensorMap &mb, const GemmArgs &g, cudaStream_t st) {
  constexpr size_t smem = (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + (2 * STAGES + 4) * 8 + 16 + 8 * STG_BYTES + 2 * BN * 4;
  auto kern = gemm_kernel<BN>;
  static int sms_of_dev[64] = {0};  // also marks "attribute set on this device" (one-time host work per device)
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return SV_ERR_INVALID_ARG;
  if (sms_of_dev[dev] == 0) {
    int rc = sv::cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (rc) return rc;
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sms_of_dev[dev] = n;
  }
  const int sms = sms_of_dev[dev];
  const int tiles = ((g.M + BM - 1) / BM) * ((g.N + BN - 1) / BN);
  const int grid = tiles < sms ? tiles : sms;
  kern<<<grid, 320, smem, st>>>(ma, mb, g);
  return sv::after_launch();
}

}  // namespace

extern "C" int sv_gemm_bf16(const void *A, int lda, const void *B, int ldb, int M, int N, int K, const float *bias,
                            int act, const void *residual, void *out, int ldo, int out_f32, int rowmax, void *stream) {
  return sv_gemm_bf16_ex(A, lda, 0, B, ldb, 0, M, N, K, bias, act, residual, out, ldo, out_f32, rowmax, stream);
}

extern "C" int sv_gemm_bf16_ex(const void *A, int lda, int a_transposed, const void *B, int ldb, int b_transposed, int M,
                               int N, int K, const float *bias, int act, const void *residual, void *out, int ldo,
                               int out_f32, int rowmax, void *stream) {
  if (M < 0 || N < 0 || K < 0) return SV_ERR_INVALID_ARG;
  if (M == 0 || N == 0) return SV_OK;
  if (!A || !B || !out || K < 8 || (lda % 8) || (ldb % 8)) return SV_ERR_INVALID_ARG;
  if (a_transposed ? lda < M : (lda < K || (K % 8))) return SV_ERR_INVALID_ARG;
  if (b_transposed ? ldb < N : (ldb < K || (K % 8))) return SV_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return SV_ERR_INVALID_ARG;
  if (act < 0 || act > 2 || (rowmax != 0 && rowmax != 16) || (rowmax && residual)) return SV_ERR_INVALID_ARG;
  if (rowmax && (M % 16)) return SV_ERR_INVALID_ARG;
  GemmArgs g{M, N, K, bias, residual, out, act, out_f32, rowmax, ldo, a_transposed ? 1 : 0, b_transposed ? 1 : 0};
  CUtensorMap ma, mb;
  const int bn = N > 128 ? 256 : (N > 64 ? 128 : 64);
  int rc = a_transposed ? make_map_mn(&ma, A, M, K, lda) : make_map(&ma, A, M, K, lda, BM);
  if (rc) return rc;
  rc = b_transposed ? make_map_mn(&mb, B, N, K, ldb) : make_map(&mb, B, N, K, ldb, bn);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (bn == 256) return launch_gemm<256>(ma, mb, g, st);
  if (bn == 128) return launch_gemm<128>(ma, mb, g, st);
  return launch_gemm<64>(ma, mb, g, st);
}
