// Persistent, warp-specialised tcgen05 GEMM with fused epilogues — the dense-contraction workhorse of the GPS path:
// every nn.Linear of the attention stack / BERT / heads in all three directions (forward, dgrad, wgrad) and the
// SA3 + fc stage of PointNet++ (reference: F.linear / 1x1 Conv2d call sites and their autograd, SURVEY.md §2.3,
// e.g. modules/layers/transformers.py:115-154,188-192,285-316).
//
//   C[M,N] = epilogue( A[M,K] x B[N,K]^T ),  bf16 operands, fp32 accumulation in TMEM
//
// Roles (448 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..13 = epilogue (three warps per TMEM lane quadrant, each thread one accumulator row, the 32-column chunks of the
// tile dealt round-robin to the three).  With K = 768 (most linears of the stack) a tile's main loop is only 12 K steps,
// so the epilogue — TMEM load, bias / activation math, staged coalesced stores — is what bounds the kernel: ncu showed the
// 8-warp version at 94 % busy epilogue warps and 57 % tensor pipe (profiles/r2_gemm_ncu_summary.json).
// Operand tiles travel global -> shared by TMA (cp.async.bulk.tensor.2d, 128-byte rows, SWIZZLE_128B) through a 4-stage
// mbarrier ring; either operand may be stored transposed in memory ([K][rows]) and is then staged as 64-column slabs and
// read MN-major by the tensor core — the dgrad / wgrad forms of a linear layer need no transposed copy.  Accumulators
// are double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.  Out-of-bounds rows / columns /
// K are zero-filled by TMA, the epilogue masks its stores.
//
// Epilogue families (template parameter EPI):
//   EPI_FWD   + bias -> (store pre-activation copy) -> relu | gelu(erf) -> dropout (counter hash, nothing stored)
//             -> + residual -> bf16 | f32 store, or max over 16-row groups (PointNet++ SA3 neighbourhood max)
//   EPI_DGRAD x activation derivative: relu+dropout from the saved forward OUTPUT (h > 0 ? 1/keep : 0), or gelu' from the
//             saved pre-activation with the dropout mask regenerated from (seed, row, column)
//   EPI_WGRAD fp32 result, split-K over the work list, `red.global.add.v4.f32` straight into the (flat) gradient buffer;
//             the bias gradient (column sums of dL/dy) comes out of the SAME main loop: one extra N = 16 MMA per K step
//             against a constant tile of ones, accumulated in 32 spare TMEM columns (first column-tile only)
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>

#include "attn_common.cuh"
#include "svcommon.h"
#include "svgps.h"
#include "tc05.cuh"

namespace {

using namespace tc05;

constexpr int BM = 128, BK = 64, STAGES = 4;
constexpr int EPI_WARPS = 12, EPI_PARTS = 3, NTHREADS = 64 + 32 * EPI_WARPS;
constexpr int STG_BYTES = 32 * 80;  // per epilogue warp: 32 rows x (64 B + 16 B pad); also holds the [32][17] f32 row-max scratch
constexpr int EPI_FWD = 0, EPI_DGRAD = 1, EPI_WGRAD = 2;

struct GemmArgs {
  int M, N, K;
  int dbg = 0;              // profiling switches (sv_gemm_force_ctas bits 8..): 1 = epilogue skips its body, 2 = producer skips TMA
  int tma_out = 0;          // bit 0: `out` leaves through TMA tile stores (mapO), bit 1: `out2` too (mapO2) — bf16, 16-byte aligned rows
  long long *prof = nullptr;  // profiling (sv_gemm_profile): [gridDim.x][8] cycle counters of the role threads, or null
  const float *bias;        // [N] or null (FWD)
  const void *residual;     // [M,N] same dtype as out, or null (FWD; added after activation / dropout)
  void *out;                // [M,N] (or [M/rowmax,N] when rowmax > 0)
  int act;                  // FWD: 0 none, 1 relu, 2 gelu(erf);  DGRAD: 0 none, 1 relu(+dropout) from output, 2 gelu(+dropout) from pre-activation
  int out_f32;              // 0 bf16, 1 f32
  int rowmax;               // 0, or 16: max over groups of 16 consecutive rows (SA3 neighbourhood max)
  int ldo;                  // leading dimension of out / residual / out2 (elements)
  int a_mn, b_mn;           // 1: the operand is stored [K][rows] (row-major, rows contiguous) = MN-major for the tensor core
  void *out2;               // FWD: bf16 [M,N] (ld = ldo) copy of the pre-activation (acc + bias), or null
  const void *aux;          // DGRAD: bf16 [M,N] saved forward tensor (act 1: output, act 2: pre-activation)
  int ld_aux;
  uint32_t t16;             // dropout threshold (0 = no dropout): keep iff 16-bit uniform >= t16
  float inv_keep;
  unsigned long long seed;
  const unsigned long long *seed_offset;
  int n_fast;               // tile order of the work list: 1 = column tiles vary fastest (A panels stream once, B stays in
                            // L2), 0 = row tiles vary fastest (the reverse); chosen by operand size
  int splits;               // split-K factor (>= 1); > 1 only with red_out
  int red_out;              // WGRAD: accumulate into out with red.global.add instead of storing
  float *bias_grad;         // WGRAD: [M] += row sums of A (the bias gradient), or null
};

// (MN-major SWIZZLE_128B operands are staged as 64-column slabs [BK rows][128 B], one TMA box each: 8-row groups 1024 B apart
// (stride byte offset), consecutive 64-element MN blocks one slab (BK * 128 B) apart (leading byte offset); a K step of 16
// rows advances the start address by 2048 B — tc05.cuh desc_lo_sw128; the attention kernels read P.V / dS.K the same way.)
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// Phi(x) = 0.5 (1 + erf(x / sqrt 2)) and e = exp(-x^2 / 2) with two MUFU ops: erf by Abramowitz-Stegun 7.1.26
// (|error| < 1.5e-7, far below the bf16 / fp32-accumulation noise of the GEMM it follows); gelu = x Phi,
// gelu' = Phi + x e / sqrt(2 pi) — the exact (erf) GELU the reference uses (F.gelu default, HF "gelu").
__device__ __forceinline__ void gelu_parts(float x, float &cdf, float &e) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  const float p = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
  e = attn::ex2f(-z * z * 1.4426950408889634f);
  const float h = 0.5f * p * e;          // = 0.5 (1 - erf|z|)
  cdf = x >= 0.f ? 1.0f - h : h;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float c, e;
  gelu_parts(x, c, e);
  return x * c;
}
__device__ __forceinline__ float gelu_grad(float x) {
  float c, e;
  gelu_parts(x, c, e);
  return fmaf(x * e, 0.3989422804014327f, c);
}

// ---- chunk stores: a warp's 32 rows x 32 columns leave through a per-warp [32][80 B] staging buffer so that one store
// instruction writes 8 rows x 64 B instead of 32 scattered 16-byte pieces ---------------------------------------------------
template <bool RED>
__device__ __forceinline__ void store_chunk_f32(const float (&v)[32], float *ob, int ldo, const float *res, int M, int N,
                                                int row, int wrow0, int col0, uint8_t *stg, int lane) {
  const bool vec_ok = col0 + 32 <= N && res == nullptr && (reinterpret_cast<uintptr_t>(ob) & 15) == 0 && (ldo & 3) == 0 &&
                      (col0 & 3) == 0;
  if (vec_ok) {
    const int sr = lane >> 2, seg = lane & 3;
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
        if (wrow0 + r < M) {
          float *dst = ob + (size_t)(wrow0 + r) * ldo + col0 + hh * 16 + seg * 4;
          if (RED) red_add_v4(dst, w.x, w.y, w.z, w.w);
          else *reinterpret_cast<float4 *>(dst) = w;
        }
      }
      __syncwarp();
    }
  } else if (row < M) {
    float *o = ob + (size_t)row * ldo + col0;
    const float *r = res ? res + (size_t)row * ldo + col0 : nullptr;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (col0 + i < N) {
        if (RED) atomicAdd(o + i, v[i]);
        else o[i] = v[i] + (r ? r[i] : 0.f);
      }
  }
}
__device__ __forceinline__ void store_chunk_bf16(const float (&v)[32], __nv_bfloat16 *ob, int ldo, const __nv_bfloat16 *res,
                                                 int M, int N, int row, int wrow0, int col0, uint8_t *stg, int lane) {
  const bool vec_ok = col0 + 32 <= N && res == nullptr && (reinterpret_cast<uintptr_t>(ob) & 15) == 0 && (ldo & 7) == 0 &&
                      (col0 & 7) == 0;
  if (vec_ok) {
    const int sr = lane >> 2, seg = lane & 3;
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
      if (wrow0 + r < M) *reinterpret_cast<uint4 *>(ob + (size_t)(wrow0 + r) * ldo + col0 + seg * 8) = w;
    }
    __syncwarp();
  } else if (row < M) {
    __nv_bfloat16 *o = ob + (size_t)row * ldo + col0;
    const __nv_bfloat16 *r = res ? res + (size_t)row * ldo + col0 : nullptr;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (col0 + i < N) o[i] = __float2bfloat16_rn(v[i] + (r ? __bfloat162float(r[i]) : 0.f));
  }
}

// ---- TMA tile stores: a warp's 32 x 32 bf16 chunk is written to its staging buffer as [32 rows][64 B] in the
// SWIZZLE_64B pattern (16-byte piece c of row r lives at piece c ^ ((r >> 1) & 3): conflict-free 16-byte stores with one
// row per lane), then ONE lane hands the 2 KB box to the TMA unit, which clips rows >= M / columns >= N itself.  Replaces
// the per-lane ld.shared + address arithmetic + predicated st.global.v4 of store_chunk_bf16: the epilogue warps were
// latency-bound at ~300 instructions per chunk (profiles: r2 gemm decomposition). ------------------------------------------
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_chunk_bf16(const float (&v)[32], const CUtensorMap *map, int col0, int wrow0, uint8_t *stg,
                                                     int lane) {
  if (lane == 0) tma_store_wait_read();   // the previous box of this warp has been read out of the staging buffer
  __syncwarp();
  const uint32_t sw = (uint32_t)(lane >> 1) & 3u;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    *reinterpret_cast<uint4 *>(stg + lane * 64 + ((k ^ sw) << 4)) =
        make_uint4(pack_bf16(v[8 * k], v[8 * k + 1]), pack_bf16(v[8 * k + 2], v[8 * k + 3]),
                   pack_bf16(v[8 * k + 4], v[8 * k + 5]), pack_bf16(v[8 * k + 6], v[8 * k + 7]));
  fence_proxy_async_smem();   // generic-proxy writes -> visible to the async proxy (every writer fences, then the warp syncs)
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(col0), "r"(wrow0), "r"(smem_u32(stg))
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}

// ---- CTA-pair plumbing (cta_group::2): the two CTAs of a cluster sit on the two SMs of one TPC and execute ONE MMA of
// M = 256 rows; each CTA stages its own 128 rows of A and HALF of the B tile, so per-CTA operand traffic through
// TMA / L2 (the ~6.3 KB/clk chip-wide limit that bounds the 128 x 256 single-CTA tile at ~1.05 PFLOP/s) drops by a third
// to a half.  Only the leader (rank 0) issues MMAs; completion is multicast to the same barrier in both CTAs. -------------
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// warp-converged forms (every lane executes, elect.sync picks the issuing lane): TMA tile load crediting `bar_addr` (a
// shared::cluster address — the own CTA's barrier, or with CTA pairs the pair leader's) and the expect_tx arrival
template <int CTAS>
__device__ __forceinline__ void tma_load_2d_elect(uint32_t dst_smem, const CUtensorMap *map, int c0, int c1, uint32_t bar_addr) {
  if (CTAS == 1) {
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n\t}"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar_addr)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n\t}"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar_addr)
        : "memory");
  }
}
// CL = 2: one box, delivered to the same shared-memory offset of every CTA in `cta_mask`; completion is signalled on the
// barrier at `bar_addr`'s offset in the LEADER of each destination CTA's pair (cta_group::2 barrier addressing)
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst_smem, const CUtensorMap *map, int c0, int c1, uint32_t bar_addr,
                                               uint16_t cta_mask) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%2, %3}], [%4], %5;\n\t}"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar_addr), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void expect_tx_elect(uint32_t bar_smem_addr, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar_smem_addr), "r"(bytes)
      : "memory");
}
template <int CTAS, uint32_t NCOLS>
__device__ __forceinline__ void tmem_alloc_g(uint32_t *smem_dst) {
  if (CTAS == 1) {
    tmem_alloc<NCOLS>(smem_dst);
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CTAS, uint32_t NCOLS>
__device__ __forceinline__ void tmem_dealloc_g(uint32_t taddr) {
  if (CTAS == 1) tmem_dealloc<NCOLS>(taddr);
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

template <int BN, int CTAS>
struct GemmCfg {
  static constexpr int BNL = BN / CTAS;                         // B rows staged by one CTA
  static constexpr int A_STAGE = BM * BK * 2, B_STAGE = BNL * BK * 2;
  static constexpr int STAGES_RAW = (192 * 1024) / (A_STAGE + B_STAGE);
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr size_t SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) + 2048 + (2 * STAGES + 4) * 8 + 16 + EPI_WARPS * STG_BYTES + 2 * BN * 4;
};

template <int BN, int EPI, int CTAS, int CL>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
            const __grid_constant__ CUtensorMap mapO, const __grid_constant__ CUtensorMap mapO2, const GemmArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using Cfg = GemmCfg<BN, CTAS>;
  constexpr int BNL = Cfg::BNL, A_STAGE = Cfg::A_STAGE, B_STAGE = Cfg::B_STAGE, STAGES = Cfg::STAGES;
  // a weight gradient keeps the bias-gradient sums in 32 spare columns per accumulator; at BN = 256 that leaves room for
  // one accumulator only (the K loop of a weight gradient is long, the un-overlapped epilogue is a few percent)
  constexpr int ACC_BUFS = (EPI == EPI_WGRAD && BN == 256) ? 1 : 2;
  constexpr uint32_t ACC_COLS = ACC_BUFS * BN;
  constexpr uint32_t NEED_COLS = EPI == EPI_WGRAD ? ACC_COLS + 32 * ACC_BUFS : ACC_COLS;
  constexpr uint32_t TMEM_COLS = NEED_COLS <= 128 ? 128 : NEED_COLS <= 256 ? 256 : 512;
  static_assert(NEED_COLS <= 512, "TMEM budget");
  static_assert(BNL % 8 == 0, "B half-tile rows");
  uint8_t *sA = smem;
  uint8_t *sB = smem + STAGES * A_STAGE;
  uint8_t *sOnes = sB + STAGES * B_STAGE;                      // [16][64] bf16 ones (2 KB), WGRAD bias-gradient operand
  float *stage = reinterpret_cast<float *>(sOnes + 2048);   // [EPI_WARPS][STG_BYTES] store staging (512-byte aligned: TMA store
                                                            // boxes, SWIZZLE_64B) / row-max transpose scratch
  uint64_t *full = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(stage) + EPI_WARPS * STG_BYTES);
  uint64_t *empty = full + STAGES;
  uint64_t *acc_full = empty + STAGES;   // [2]
  uint64_t *acc_empty = acc_full + 2;    // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
  float *sbias = reinterpret_cast<float *>(tmem_slot + 4);      // [2][BN] bias slice of the tile, double-buffered

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CL = 2 (CTA pairs only): a cluster of TWO pairs works on two vertically adjacent tiles of the same column block; the
  // B tile they share is fetched once — every CTA loads 1/CL of its pair-half and the TMA unit multicasts it to the CTA of
  // the same rank in the other pair — so a CTA pulls 24 KB instead of 32 KB per k-step through L2.
  static_assert(CL == 1 || CTAS == 2, "multicast clusters are built from CTA pairs");
  const uint32_t crank = CTAS == 2 ? cluster_rank() : 0u;   // rank in the cluster (0..CTAS * CL - 1)
  const uint32_t rank = crank & 1u;                          // rank in the pair
  const uint32_t pair = crank >> 1;                          // pair in the cluster
  const uint32_t leader = crank & ~1u;                       // cluster rank of this pair's leader
  const int unit = blockIdx.x / (CTAS * CL), n_units = gridDim.x / (CTAS * CL);   // a unit = one CTA, one pair, or CL pairs
  const int tiles_m1 = (g.M + CTAS * BM - 1) / (CTAS * BM);                       // row tiles of one CTA / pair
  const int tiles_m = (tiles_m1 + CL - 1) / CL, tiles_n = (g.N + BN - 1) / BN;    // row tiles of a unit
  const int n_tiles = tiles_m * tiles_n;
  const int k_steps = (g.K + BK - 1) / BK;
  const int kps = (k_steps + g.splits - 1) / g.splits;   // K steps per split (host guarantees every split is non-empty)
  const int n_items = n_tiles * g.splits;                // work list: item i = (tile i % n_tiles, split i / n_tiles)
  const bool with_bias_grad = EPI == EPI_WGRAD && g.bias_grad != nullptr;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, CL);   // one commit per pair of the cluster
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full + b, 1);
      mbar_init(acc_empty + b, CTAS * EPI_WARPS);    // one arrival per epilogue warp of every CTA of the unit
    }
    mbar_fence_init();
  }
  if (EPI == EPI_WGRAD) {
    for (int i = threadIdx.x; i < 512; i += NTHREADS) reinterpret_cast<uint32_t *>(sOnes)[i] = 0x3F803F80u;  // bf16 1.0 pairs
    fence_proxy_async_smem();
  }
  if (warp == 1) tmem_alloc_g<CTAS, TMEM_COLS>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();     // the peer's barriers are initialised before anything remote touches them
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------- TMA producer (every CTA stages its own rows of A and its share of B) -------------
    // The whole warp runs the loop converged; only the TMA / expect_tx instructions are predicated on elect.sync, so the
    // coordinates and addresses stay warp-uniform (see tc05.cuh: warp-converged issue).
    {
      const uint32_t full0 = smem_u32(full), sA0 = smem_u32(sA), sB0 = smem_u32(sB);
      const uint32_t fb0 = CTAS == 2 ? map_to_rank(full0, leader) : full0;   // the pair leader's barriers
      const uint16_t bmask = (uint16_t)((1u << rank) | (1u << (rank + 2)));      // CL = 2: same rank in both pairs
      uint32_t s = 0, ph = 1;   // waiting on parity 1 of a fresh barrier returns at once: the first pass finds every slot free
      for (int item = unit; item < n_items; item += n_units) {
        const int t = item % n_tiles, sp = item / n_tiles;
        const int tmi = g.n_fast ? t / tiles_n : t % tiles_m, tni = g.n_fast ? t % tiles_n : t / tiles_m;
        const int m0 = ((tmi * CL + (int)pair) * CTAS + (int)rank) * BM, n0 = tni * BN + (int)rank * BNL;
        const int ks0 = sp * kps, ks1 = min(k_steps, ks0 + kps);
        for (int ks = ks0; ks < ks1; ++ks) {
          mbar_wait(empty + s, ph);  // slot free (in every CTA this CTA's loads land in)
          const int k0 = ks * BK;
          const uint32_t da = sA0 + s * A_STAGE, db = sB0 + s * B_STAGE, fb = fb0 + s * 8;
          if (g.dbg & 2) {   // profiling: no loads, the MMAs chew on whatever the slot holds
            if (rank == 0 && lane == 0) mbar_arrive(full + s);
          } else {
            if (CTAS == 1) expect_tx_elect(fb, A_STAGE + B_STAGE);
            else if (rank == 0) expect_tx_elect(full0 + s * 8, 2 * (A_STAGE + B_STAGE));  // both CTAs' bytes land on the leader's barrier
            if (g.a_mn) {  // [K][M] storage: one [BK x 64] box per 64 rows of the tile
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_2d_elect<CTAS>(da + j * (BK * 128), &mapA, m0 + j * 64, k0, fb);
            } else {
              tma_load_2d_elect<CTAS>(da, &mapA, k0, m0, fb);
            }
            if (CL == 2) {
              // this CTA's share of the pair-half: BNL / 2 rows, multicast to the same-rank CTA of both pairs
              constexpr int QR = BNL / 2;
              if (g.b_mn) {
#pragma unroll
                for (int j = 0; j < QR / 64; ++j)
                  tma_load_2d_mc(db + ((int)pair * (QR / 64) + j) * (BK * 128), &mapB, n0 + (int)pair * QR + j * 64, k0, fb, bmask);
              } else {
                tma_load_2d_mc(db + (int)pair * QR * 128, &mapB, k0, n0 + (int)pair * QR, fb, bmask);
              }
            } else if (g.b_mn) {
#pragma unroll
              for (int j = 0; j < BNL / 64; ++j) tma_load_2d_elect<CTAS>(db + j * (BK * 128), &mapB, n0 + j * 64, k0, fb);
            } else {
              tma_load_2d_elect<CTAS>(db, &mapB, k0, n0, fb);
            }
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (leader CTA only) -------------------------------
    // Warp-converged: all 32 lanes run the loop, elect.sync picks the lane that issues (tc05.cuh).  Per MMA the loop body
    // is two 32-bit adds on the descriptors' low words — the single-lane version spent 580-760 cycles per k-step (four
    // 128-cycle MMAs) on descriptor arithmetic and the compiler's per-instruction ELECT loops and bounded the whole kernel.
    if (rank == 0) {
      const uint32_t IDESC = make_idesc_bf16(CTAS * BM, BN) | (g.a_mn ? 1u << 15 : 0u) | (g.b_mn ? 1u << 16 : 0u);
      const uint32_t IDESC_ONES = make_idesc_bf16(CTAS * BM, 16) | (g.a_mn ? 1u << 15 : 0u);   // B = K-major tile of ones
      const uint32_t a_lo0 = desc_lo_sw128(smem_u32(sA), g.a_mn != 0, BK * 128), b_lo0 = desc_lo_sw128(smem_u32(sB), g.b_mn != 0, BK * 128);
      const uint32_t a_step = g.a_mn ? 2048u >> 4 : 32u >> 4, b_step = g.b_mn ? 2048u >> 4 : 32u >> 4;   // one K = 16 step
      const uint32_t ones_lo = desc_lo_sw128(smem_u32(sOnes), false, 0);
      const uint32_t empty0 = smem_u32(empty), accfull0 = smem_u32(acc_full);
      uint32_t s = 0, ph = 0, tl = 0, nks = 0;
      long long p_full = 0, p_acc = 0, p_t0 = g.prof ? clock64() : 0;
      unsigned long long p_ns0 = 0;
      if (g.prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(p_ns0));
      for (int item = unit; item < n_items; item += n_units, ++tl) {
        const int t = item % n_tiles, sp = item / n_tiles;
        const bool bias_tile = with_bias_grad && (g.n_fast ? t % tiles_n : t / tiles_m) == 0;
        const int ks0 = sp * kps, ks1 = min(k_steps, ks0 + kps);
        const int b = tl % ACC_BUFS;
        long long c0 = g.prof ? clock64() : 0;
        mbar_wait(acc_empty + b, ((tl / ACC_BUFS) & 1u) ^ 1u);  // every epilogue warp of the unit drained this accumulator
        fence_after_sync();
        if (g.prof) p_acc += clock64() - c0;
        const uint32_t d_acc = tmem + b * BN, d_bias = tmem + ACC_COLS + b * 32;
        uint32_t acc = 0;
        for (int ks = ks0; ks < ks1; ++ks, ++nks) {
          if (g.prof) c0 = clock64();
          mbar_wait(full + s, ph);
          fence_after_sync();
          if (g.prof) p_full += clock64() - c0;
          const uint32_t al = a_lo0 + s * (A_STAGE >> 4), bl = b_lo0 + s * (B_STAGE >> 4);
          if (EPI == EPI_WGRAD && bias_tile) {
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
              mma_bf16_elect<CTAS>(d_acc, al + kk * a_step, DESC_HI_SW128, bl + kk * b_step, DESC_HI_SW128, IDESC, acc | kk);
              mma_bf16_elect<CTAS>(d_bias, al + kk * a_step, DESC_HI_SW128, ones_lo, DESC_HI_SW128, IDESC_ONES, acc | kk);
            }
          } else {
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk)
              mma_bf16_elect<CTAS>(d_acc, al + kk * a_step, DESC_HI_SW128, bl + kk * b_step, DESC_HI_SW128, IDESC, acc | kk);
          }
          acc = 1;
          mma_commit_elect<CTAS>(empty0 + s * 8, (uint16_t)((1u << (CTAS * CL)) - 1u));  // frees the slot in every CTA of the cluster
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        mma_commit_elect<CTAS>(accfull0 + b * 8, (uint16_t)(3u << leader));   // both CTAs of THIS pair
      }
      if (g.prof && lane == 0) {   // issue-loop cycles, of which waiting for operands / for a free accumulator, k-steps, start, end
        long long *pp = g.prof + (size_t)blockIdx.x * 16;
        const long long t1 = clock64();
        unsigned long long ns1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
        pp[0] = t1 - p_t0; pp[1] = p_full; pp[2] = p_acc; pp[3] = nks; pp[4] = p_t0; pp[5] = t1; pp[6] = (long long)(ns1 - p_ns0);
      }
    }
  } else {
    // ------------------------------- epilogue (warps 2..13) -------------------------------
    // Three warps per TMEM lane quadrant; chunk c (32 columns) of the tile belongs to warp part c % 3.  Per tile the bias
    // slice is staged once in shared memory (double-buffered), TMEM loads run one chunk ahead of the math.
    const int q = warp & 3;             // TMEM lane quadrant this warp may access (hardware rule: warp id % 4)
    const int part = (warp - 2) >> 2;   // 0..2
    const int et = threadIdx.x - 64;    // 0..383
    const int row_in_tile = q * 32 + lane;
    constexpr int NCH = BN / 32;        // 32-column chunks of the tile
    uint8_t *stg = reinterpret_cast<uint8_t *>(stage) + (warp - 2) * STG_BYTES;
    const unsigned long long seed = (EPI != EPI_WGRAD && g.t16) ? attn::effective_seed(g.seed, g.seed_offset) : 0ull;
    uint32_t tl = 0;
    long long e_bar = 0, e_acc = 0, e_body = 0, e_c = 0;
    for (int item = unit; item < n_items; item += n_units, ++tl) {
      const int t = item % n_tiles;
      const int b = tl % ACC_BUFS;
      if (g.prof) e_c = clock64();
      const int tmi = g.n_fast ? t / tiles_n : t % tiles_m, tni = g.n_fast ? t % tiles_n : t / tiles_m;
      const int m0 = ((tmi * CL + (int)pair) * CTAS + (int)rank) * BM, n0 = tni * BN;
      const int row = m0 + row_in_tile;
      const int wrow0 = m0 + q * 32;  // first accumulator row of this warp
      float *sb = sbias + (tl & 1) * BN;
      if (EPI == EPI_FWD) {
        if (et < BN) sb[et] = (g.bias != nullptr && n0 + et < g.N) ? __ldg(g.bias + n0 + et) : 0.f;
        asm volatile("bar.sync 1, 384;" ::: "memory");  // bias visible; everybody is done with the tile before last
      }
      uint32_t rk = 0;
      if (EPI != EPI_WGRAD && g.t16) rk = attn::drop_row_key(seed, (unsigned long long)row);
      if (g.prof) { const long long c = clock64(); e_bar += c - e_c; e_c = c; }
      mbar_wait(acc_full + b, (tl / ACC_BUFS) & 1u);
      fence_after_sync();
      if (g.prof) { const long long c = clock64(); e_acc += c - e_c; e_c = c; }
      const uint32_t taddr = tmem + b * BN + ((uint32_t)(q * 32) << 16);
      uint32_t rr[2][32];
      if (part < NCH) tmem_ld32_async(taddr + part * 32, rr[0]);
#pragma unroll
      for (int cj = 0; cj < (NCH + EPI_PARTS - 1) / EPI_PARTS; ++cj) {
        const int ci = part + cj * EPI_PARTS;
        if (ci >= NCH) break;
        uint32_t(&cur)[32] = rr[cj & 1];
        tmem_wait32(cur);
        if (ci + EPI_PARTS < NCH) tmem_ld32_async(taddr + (ci + EPI_PARTS) * 32, rr[(cj + 1) & 1]);
        const int c0 = ci * 32;
        const int col0 = n0 + c0;
        if (col0 < g.N && !(g.dbg & 1)) {
          float v[32];
          if (EPI == EPI_FWD) {
            const float4 *b4 = reinterpret_cast<const float4 *>(sb + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 bb = b4[i];
              v[4 * i] = __uint_as_float(cur[4 * i]) + bb.x;
              v[4 * i + 1] = __uint_as_float(cur[4 * i + 1]) + bb.y;
              v[4 * i + 2] = __uint_as_float(cur[4 * i + 2]) + bb.z;
              v[4 * i + 3] = __uint_as_float(cur[4 * i + 3]) + bb.w;
            }
            if (g.out2 != nullptr) {
              if (g.tma_out & 2) tma_store_chunk_bf16(v, &mapO2, col0, wrow0, stg, lane);
              else store_chunk_bf16(v, reinterpret_cast<__nv_bfloat16 *>(g.out2), g.ldo, nullptr, g.M, g.N, row, wrow0, col0, stg, lane);
            }
            if (g.act == 1) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            } else if (g.act == 2) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
            }
            if (g.t16) {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const uint32_t hh = attn::drop_pair_hash(rk, (uint32_t)(col0 + i) >> 1);
                v[i] = (hh & 0xFFFFu) >= g.t16 ? v[i] * g.inv_keep : 0.f;
                v[i + 1] = (hh >> 16) >= g.t16 ? v[i + 1] * g.inv_keep : 0.f;
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(cur[i]);
          }
          if (EPI == EPI_DGRAD && g.act != 0 && row < g.M) {
            // multiply by the activation (+ dropout) derivative read from the saved forward tensor
            const __nv_bfloat16 *ax = reinterpret_cast<const __nv_bfloat16 *>(g.aux) + (size_t)row * g.ld_aux + col0;
            const bool vec = col0 + 32 <= g.N && (g.ld_aux & 7) == 0 && (reinterpret_cast<uintptr_t>(g.aux) & 15) == 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float a8[8];
              if (vec) {
                const uint4 w = __ldg(reinterpret_cast<const uint4 *>(ax) + k);
                const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  a8[2 * e] = __uint_as_float(ww[e] << 16);
                  a8[2 * e + 1] = __uint_as_float(ww[e] & 0xFFFF0000u);
                }
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) a8[e] = col0 + k * 8 + e < g.N ? __bfloat162float(ax[k * 8 + e]) : 0.f;
              }
              if (g.act == 1) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[k * 8 + e] = a8[e] > 0.f ? v[k * 8 + e] * g.inv_keep : 0.f;
              } else {
#pragma unroll
                for (int e = 0; e < 8; e += 2) {
                  float m0_ = 1.f, m1_ = 1.f;
                  if (g.t16) {
                    const uint32_t hh = attn::drop_pair_hash(rk, (uint32_t)(col0 + k * 8 + e) >> 1);
                    m0_ = (hh & 0xFFFFu) >= g.t16 ? g.inv_keep : 0.f;
                    m1_ = (hh >> 16) >= g.t16 ? g.inv_keep : 0.f;
                  }
                  v[k * 8 + e] *= gelu_grad(a8[e]) * m0_;
                  v[k * 8 + e + 1] *= gelu_grad(a8[e + 1]) * m1_;
                }
              }
            }
          }
          if (EPI == EPI_FWD && g.rowmax == 16) {
            // max over the 16 rows of each half-warp (16 consecutive rows = the 16 points of one cloud): 16 columns at a
            // time go through a per-warp [32][17] transpose scratch; lanes 0-15 reduce rows 0-15, lanes 16-31 rows 16-31
            float *tp = reinterpret_cast<float *>(stg);
            const int l16 = lane & 15, grp = lane >> 4;
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
          } else if (EPI == EPI_WGRAD) {
            if (g.red_out)
              store_chunk_f32<true>(v, reinterpret_cast<float *>(g.out), g.ldo, nullptr, g.M, g.N, row, wrow0, col0, stg, lane);
            else
              store_chunk_f32<false>(v, reinterpret_cast<float *>(g.out), g.ldo, nullptr, g.M, g.N, row, wrow0, col0, stg, lane);
          } else if (g.out_f32) {
            store_chunk_f32<false>(v, reinterpret_cast<float *>(g.out), g.ldo, reinterpret_cast<const float *>(g.residual), g.M,
                                   g.N, row, wrow0, col0, stg, lane);
          } else if (g.tma_out & 1) {
            tma_store_chunk_bf16(v, &mapO, col0, wrow0, stg, lane);
          } else {
            store_chunk_bf16(v, reinterpret_cast<__nv_bfloat16 *>(g.out), g.ldo,
                             reinterpret_cast<const __nv_bfloat16 *>(g.residual), g.M, g.N, row, wrow0, col0, stg, lane);
          }
        }
      }
      if (EPI == EPI_WGRAD && with_bias_grad && n0 == 0 && part == 0) {
        // every column of the N = 16 ones-product equals sum_k A[row][k]: the bias gradient of output feature `row`
        uint32_t r16[16];
        attn::tmem_ld16_async(tmem + ACC_COLS + b * 32 + ((uint32_t)(q * 32) << 16), r16);
        attn::tmem_wait16(r16);
        if (row < g.M) atomicAdd(g.bias_grad + row, __uint_as_float(r16[0]));
      }
      if (g.prof) e_body += clock64() - e_c;
      // this warp has finished reading the accumulator: one arrival per warp on the LEADER's barrier
      fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (CTAS == 1) mbar_arrive(acc_empty + b);
        else mbar_arrive_cluster(map_to_rank(smem_u32(acc_empty + b), leader));
      }
    }
    if (g.tma_out && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every box has landed before the CTA exits
    if (g.prof && lane == 0 && q == 2 && part < 2) {
      // epilogue warps 2 / 6 (part 0 / 1 of lane quadrant 2): cycles at the bias barrier, waiting for the accumulator, in the body
      long long *pp = g.prof + (size_t)blockIdx.x * 16 + 8 + part * 4;
      pp[0] = e_bar; pp[1] = e_acc; pp[2] = e_body; pp[3] = tl;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();     // no CTA of the pair exits (or frees TMEM) while the other may still signal it
  if (warp == 1) tmem_dealloc_g<CTAS, TMEM_COLS>(tmem);
}

using attn::EncodeTiledFn;
using attn::encode_fn;

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

// output [rows, cols] bf16 with leading dimension ld: box = one epilogue chunk, 32 rows x 32 columns (64-byte rows, SWIZZLE_64B)
int make_map_out(CUtensorMap *map, const void *ptr, int rows, int cols, int ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return SV_ERR_CUDA;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
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

int device_sms(int *dev_out) {
  static int sms_of_dev[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 0;
  if (sms_of_dev[dev] == 0) {
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sms_of_dev[dev] = n;
  }
  *dev_out = dev;
  return sms_of_dev[dev];
}

template <int BN, int EPI, int CTAS, int CL = 1>
int launch_gemm(const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &mo, const CUtensorMap &mo2, const GemmArgs &g,
                cudaStream_t st) {
  constexpr size_t smem = GemmCfg<BN, CTAS>::SMEM;
  auto kern = gemm_kernel<BN, EPI, CTAS, CL>;
  static bool configured[64] = {false};
  int dev = 0;
  const int sms = device_sms(&dev);
  if (sms == 0) return SV_ERR_INVALID_ARG;
  if (!configured[dev]) {
    int rc = sv::cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (rc) return rc;
    configured[dev] = true;
  }
  const int tm1 = (g.M + CTAS * BM - 1) / (CTAS * BM);
  const int units = ((tm1 + CL - 1) / CL) * ((g.N + BN - 1) / BN) * g.splits;
  int max_units = sms / (CTAS * CL);
  if (CL > 1) {
    // a cluster's CTAs must sit in one GPC: fewer clusters of four are co-resident than sms / 4 (a persistent kernel with more
    // clusters than fit would run the rest as a second wave)
    static int max_clusters[64] = {0};
    if (max_clusters[dev] == 0) {
      cudaLaunchConfig_t q{};
      q.gridDim = dim3(sms / (CTAS * CL) * CTAS * CL);
      q.blockDim = dim3(NTHREADS);
      q.dynamicSmemBytes = smem;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = CTAS * CL;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      q.attrs = qa;
      q.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess || n < 1) {
        cudaGetLastError();
        n = max_units;
      }
      max_clusters[dev] = n;
      if (getenv("SV_GEMM_VERBOSE")) fprintf(stderr, "[svgps] co-resident clusters of %d CTAs: %d\n", CTAS * CL, n);
    }
    if (max_clusters[dev] < max_units) max_units = max_clusters[dev];
  }
  const int grid = (units < max_units ? units : max_units) * CTAS * CL;
  if (CTAS == 1) {
    kern<<<grid, NTHREADS, smem, st>>>(ma, mb, mo, mo2, g);
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTAS * CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ma, mb, mo, mo2, g);
    if (e != cudaSuccess) {
      sv::t_last_cuda_error = (int)e;
      cudaGetLastError();
      return SV_ERR_CUDA;
    }
  }
  return sv::after_launch();
}

// Tile selection.  ncu + the layout micro-benchmark (profiles/r2_gemm_ncu_summary.json, r2_mma_bench.json) show the main loop
// bounded by SHARED-MEMORY traffic, not by the tensor pipe or L2: every K step the TMA writes the stage (A: 16 KB, B:
// 128 B x the rows this CTA stages) and the tensor core reads it back (A: 16 KB, B: all BN rows, half of them from the peer
// CTA in pair mode) at ~108 B/clk combined, against 2 x BN clk of math.  Cost of one tile = max(math, traffic / 108); the
// persistent grid runs ceil(tiles / units) waves of it.  `mn_b`: B is a transposed (MN-major) operand staged in 64-column
// slabs, so a CTA pair needs BN / 2 % 64 == 0.
float tile_cost(int bn, int ctas) {
  const float math = 2.0f * bn;
  const float wr = 16384.f + 128.f * bn / ctas, rd = 16384.f + 128.f * bn;
  const float mem = (wr + rd) / 108.f;
  return math > mem ? math : mem;
}

void pick_tile(int M, int N, int sms, bool mn_b, bool allow_pair, int force_ctas, bool wgrad, int *bn_out, int *ctas_out) {
  const int cand[4] = {256, 192, 128, 64};
  float best_cost = 1e30f;
  *bn_out = 64;
  *ctas_out = 1;
  for (int ctas = 1; ctas <= 2; ++ctas) {
    if (force_ctas && ctas != force_ctas && !(force_ctas == 2 && M <= BM && ctas == 1)) continue;
    if (ctas == 2 && (!allow_pair || M <= BM)) continue;
    const int units = sms / ctas;
    const int tm = (M + ctas * BM - 1) / (ctas * BM);
    for (int i = 0; i < 4; ++i) {
      const int bn = cand[i];
      if (ctas == 2 && (bn == 64 || (mn_b && (bn / 2) % 64))) continue;
      if (bn > 64 && bn - 64 >= ((N + 63) / 64) * 64) continue;          // wider than the problem by a whole 64-column slab
      const int tiles = tm * ((N + bn - 1) / bn);
      // a weight gradient fills the machine through split-K, so only the per-area cost of the tile matters there
      const float waves = wgrad ? (float)tiles / units : (float)((tiles + units - 1) / units);
      const float cost = waves * tile_cost(bn, ctas);
      if (cost < best_cost * 0.999f) {
        best_cost = cost;
        *bn_out = bn;
        *ctas_out = ctas;
      }
    }
  }
}

template <int EPI, int CTAS>
int dispatch(int bn, int cl, const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &mo, const CUtensorMap &mo2, const GemmArgs &g,
             cudaStream_t st) {
  if constexpr (CTAS == 2) {
    if (cl == 2 && bn == 256) return launch_gemm<256, EPI, 2, 2>(ma, mb, mo, mo2, g, st);
  }
  if (bn == 256) return launch_gemm<256, EPI, CTAS>(ma, mb, mo, mo2, g, st);
  if (bn == 192) return launch_gemm<192, EPI, CTAS>(ma, mb, mo, mo2, g, st);
  if (bn == 128) return launch_gemm<128, EPI, CTAS>(ma, mb, mo, mo2, g, st);
  if constexpr (CTAS == 1) return launch_gemm<64, EPI, 1>(ma, mb, mo, mo2, g, st);
  return SV_ERR_INVALID_ARG;
}

int g_force_ctas = 0;   // tests / benchmarks: 1 or 2 forces the single-CTA or the CTA-pair kernel, 4 = pairs in multicast clusters
                        // of two (256-wide tiles), 0 = heuristic
int g_gemm_dbg = 0;     // profiling only (bits 8.. of sv_gemm_force_ctas): results are garbage when set
long long *g_gemm_prof = nullptr;

int run_gemm(int epi, const void *A, int lda, int a_t, const void *B, int ldb, int b_t, GemmArgs &g, cudaStream_t st) {
  const int M = g.M, N = g.N, K = g.K;
  if (M < 0 || N < 0 || K < 0) return SV_ERR_INVALID_ARG;
  if (M == 0 || N == 0) return SV_OK;
  if (!A || !B || !g.out || K < 1 || (lda % 8) || (ldb % 8)) return SV_ERR_INVALID_ARG;
  if (a_t ? lda < M : lda < K) return SV_ERR_INVALID_ARG;   // K itself may be anything: TMA zero-fills past the extent
  if (b_t ? ldb < N : ldb < K) return SV_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return SV_ERR_INVALID_ARG;
  int dev = 0;
  const int sms = device_sms(&dev);
  if (sms == 0) return SV_ERR_INVALID_ARG;
  g.a_mn = a_t ? 1 : 0;
  g.b_mn = b_t ? 1 : 0;
  // the operand with the SMALLER footprint is the one re-read across the in-flight tiles (it stays in L2); the larger one
  // is streamed from HBM exactly once
  g.n_fast = (long long)N <= (long long)M ? 1 : 0;
  g.dbg = g_gemm_dbg;
  g.prof = g_gemm_prof;
  int bn = 64, ctas = 1;
  pick_tile(M, N, sms, b_t != 0, g.rowmax == 0, g_force_ctas == 4 ? 2 : g_force_ctas, epi == EPI_WGRAD, &bn, &ctas);
  // two pairs per cluster sharing the B tile by TMA multicast: 256-wide pair tiles with at least two row tiles
  int cl = 1;
  if (ctas == 2 && bn == 256 && M > 2 * BM && g_force_ctas == 4 && epi != EPI_WGRAD) cl = 2;
  const int units = sms / (ctas * cl);
  g.splits = 1;
  if (epi == EPI_WGRAD && g.red_out) {
    // split-K: the output of a weight gradient is small (a few dozen tiles), the contraction runs over every token
    const int tm1 = (M + ctas * BM - 1) / (ctas * BM);
    const int tiles = ((tm1 + cl - 1) / cl) * ((N + bn - 1) / bn), k_steps = (K + BK - 1) / BK;
    float best_eff = 0.f;
    for (int s = 1; s <= 32 && s * 4 <= k_steps; ++s) {
      const int kps = (k_steps + s - 1) / s;
      if ((s - 1) * kps >= k_steps) continue;  // an empty split
      const int items = tiles * s;
      const float eff = (float)items / (float)(((items + units - 1) / units) * units) - 0.005f * s;
      if (eff > best_eff) {
        best_eff = eff;
        g.splits = s;
      }
    }
  }
  CUtensorMap ma, mb;
  int rc = a_t ? make_map_mn(&ma, A, M, K, lda) : make_map(&ma, A, M, K, lda, BM);
  if (rc) return rc;
  rc = b_t ? make_map_mn(&mb, B, N, K, ldb) : make_map(&mb, B, N, K, ldb, bn / ctas / cl);
  if (rc) return rc;
  // bf16 outputs with 16-byte aligned rows leave through TMA tile stores (the row-max and residual epilogues keep the
  // per-lane path: they do not store the chunk as it is)
  CUtensorMap mo = ma, mo2 = ma;   // placeholders when unused (never dereferenced by the kernel)
  g.tma_out = 0;
  auto tma_ok = [&](const void *ptr) {
    return ptr != nullptr && !g.out_f32 && (g.ldo % 8) == 0 && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0;
  };
  if (epi != EPI_WGRAD && g.rowmax == 0 && g.residual == nullptr && tma_ok(g.out)) {
    rc = make_map_out(&mo, g.out, M, N, g.ldo);
    if (rc) return rc;
    g.tma_out |= 1;
  }
  if (epi == EPI_FWD && tma_ok(g.out2)) {
    rc = make_map_out(&mo2, g.out2, M, N, g.ldo);
    if (rc) return rc;
    g.tma_out |= 2;
  }
  if (ctas == 2) {
    if (epi == EPI_FWD) return dispatch<EPI_FWD, 2>(bn, cl, ma, mb, mo, mo2, g, st);
    if (epi == EPI_DGRAD) return dispatch<EPI_DGRAD, 2>(bn, cl, ma, mb, mo, mo2, g, st);
    return dispatch<EPI_WGRAD, 2>(bn, cl, ma, mb, mo, mo2, g, st);
  }
  if (epi == EPI_FWD) return dispatch<EPI_FWD, 1>(bn, 1, ma, mb, mo, mo2, g, st);
  if (epi == EPI_DGRAD) return dispatch<EPI_DGRAD, 1>(bn, 1, ma, mb, mo, mo2, g, st);
  return dispatch<EPI_WGRAD, 1>(bn, 1, ma, mb, mo, mo2, g, st);
}

void set_dropout(GemmArgs &g, float p, unsigned long long seed) {
  g.t16 = attn::drop_threshold(p);
  g.inv_keep = 1.0f / (1.0f - p);
  g.seed = seed;
  g.seed_offset = sv::g_seed_offset;
}

}  // namespace

extern "C" int sv_gemm_profile(long long *buf) {
  g_gemm_prof = buf;
  return SV_OK;
}

extern "C" int sv_gemm_force_ctas(int ctas) {
  g_force_ctas = ctas & 0xFF;
  g_gemm_dbg = ctas >> 8;
  return SV_OK;
}

extern "C" int sv_gemm_bf16(const void *A, int lda, const void *B, int ldb, int M, int N, int K, const float *bias,
                            int act, const void *residual, void *out, int ldo, int out_f32, int rowmax, void *stream) {
  return sv_gemm_bf16_ex(A, lda, 0, B, ldb, 0, M, N, K, bias, act, residual, out, ldo, out_f32, rowmax, stream);
}

extern "C" int sv_gemm_bf16_ex(const void *A, int lda, int a_transposed, const void *B, int ldb, int b_transposed, int M,
                               int N, int K, const float *bias, int act, const void *residual, void *out, int ldo,
                               int out_f32, int rowmax, void *stream) {
  if (act < 0 || act > 2 || (rowmax != 0 && rowmax != 16) || (rowmax && residual)) return SV_ERR_INVALID_ARG;
  if (rowmax && (M % 16)) return SV_ERR_INVALID_ARG;
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K; g.bias = bias; g.residual = residual; g.out = out; g.act = act; g.out_f32 = out_f32;
  g.rowmax = rowmax; g.ldo = ldo; g.inv_keep = 1.f;
  return run_gemm(EPI_FWD, A, lda, a_transposed, B, ldb, b_transposed, g, (cudaStream_t)stream);
}

extern "C" int sv_linear_fwd_bf16(const void *x, int ldx, const void *w, int ldw, int M, int N, int K, const float *bias,
                                  int act, float dropout_p, unsigned long long seed, void *out, int ldo, int out_f32,
                                  void *pre_out, void *stream) {
  if (act < 0 || act > 2 || !(dropout_p >= 0.f) || dropout_p >= 1.f) return SV_ERR_INVALID_ARG;
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K; g.bias = bias; g.out = out; g.act = act; g.out_f32 = out_f32; g.ldo = ldo; g.out2 = pre_out;
  set_dropout(g, dropout_p, seed);
  return run_gemm(EPI_FWD, x, ldx, 0, w, ldw, 0, g, (cudaStream_t)stream);
}

extern "C" int sv_linear_dgrad_bf16(const void *gy, int ldg, const void *w, int ldw, int M, int N, int Kin, int dact,
                                    const void *aux, int ld_aux, float dropout_p, unsigned long long seed, void *dx, int ldx,
                                    int out_f32, void *stream) {
  // dx[M,Kin] = gy[M,N] . w[N,Kin]: contraction over N, the weight is the transposed (MN-major) B operand
  if (dact < 0 || dact > 2 || (dact && !aux) || !(dropout_p >= 0.f) || dropout_p >= 1.f) return SV_ERR_INVALID_ARG;
  GemmArgs g{};
  g.M = M; g.N = Kin; g.K = N; g.out = dx; g.act = dact; g.out_f32 = out_f32; g.ldo = ldx; g.aux = aux; g.ld_aux = ld_aux;
  set_dropout(g, dropout_p, seed);
  return run_gemm(EPI_DGRAD, gy, ldg, 0, w, ldw, 1, g, (cudaStream_t)stream);
}

extern "C" int sv_linear_wgrad_bf16(const void *gy, int ldg, const void *x, int ldx, int M, int N, int Kin, float *dw,
                                    int ld_dw, float *db, int accumulate, void *stream) {
  // dw[N,Kin] (+)= gy[M,N]^T . x[M,Kin]: contraction over the M tokens, both operands transposed in memory; db[N] (+)= column
  // sums of gy.  accumulate = 0: dw / db are overwritten (zero-filled here first, then reduced into by the split-K work list)
  if (M < 0 || N < 0 || Kin < 0) return SV_ERR_INVALID_ARG;
  if (N == 0 || Kin == 0) return SV_OK;
  if (!dw) return SV_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) {
    int rc = sv::cuda_status(cudaMemsetAsync(dw, 0, (size_t)N * ld_dw * sizeof(float), st));
    if (rc) return rc;
    if (db) {
      rc = sv::cuda_status(cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), st));
      if (rc) return rc;
    }
  }
  if (M == 0) return SV_OK;
  GemmArgs g{};
  g.M = N; g.N = Kin; g.K = M; g.out = dw; g.out_f32 = 1; g.ldo = ld_dw; g.red_out = 1; g.bias_grad = db; g.inv_keep = 1.f;
  return run_gemm(EPI_WGRAD, gy, ldg, 1, x, ldx, 1, g, st);
}
