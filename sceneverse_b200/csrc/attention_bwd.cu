// Backward of the fused attention (csrc/attention.cu) on tcgen05, two kernels per call:
//   B1 (thread == query row):  S = Q K^T, dP = dO V^T  ->  dS = P o (dP - D)  ->  dQ = scale * dS K,
//                              gradients of the 6 spatial-gate weights of each (query, head), D = rowsum(dO o O)
//   B2 (thread == key row):    S^T = K Q^T, dP^T = V dO^T  ->  P^T, dS^T  ->  dV = P^T dO,  dK = scale * dS^T Q
// P is recomputed from Q, K, the gate and the saved log-sum-exp (no (B,H,L,T) tensor is ever stored).  Every MMA
// operand is either a row-major (K-major) tile or its explicit transpose staged once in shared memory — the same two
// operand patterns the forward kernel uses.  Reference math: modules/layers/transformers.py:188-237 (autograd of it).
#include <cuda_bf16.h>

#include "svcommon.h"
#include "svgps.h"
#include "attn_rng.cuh"
#include "tc05.cuh"

namespace {

using namespace tc05;
constexpr int DH = 64;

struct BwdArgs {
  const __nv_bfloat16 *q, *k, *v;
  long long q_bs, k_bs, v_bs;
  int q_rs, k_rs, v_rs;
  const __nv_bfloat16 *o, *d_o;      // (B,Lq,H*64) contiguous
  const unsigned char *kpm;          // (B,Lk) or null
  const float *sw, *locs, *lse;      // (B,Lq,H*6) | (B,Lq,Lk,5) | (B,H,Lq)
  int B, H, Lq, Lk;
  float scale;
  __nv_bfloat16 *dq, *dk, *dv;       // contiguous (B,L,H*64)
  float *dsw;                        // (B,Lq,H*6) or null
  float *dvec;                       // (B,H,Lq)
  unsigned drop_thresh;              // same dropout mask as the forward (0 = off)
  float inv_keep;
  unsigned long long seed;
};

// rows [r0, r0+128) of a (L, *) bf16 matrix (head slice of 64 columns) -> K-major tile [128 x 64]; rows >= L are zero
__device__ __forceinline__ void stage_tile128(uint8_t *dst, const __nv_bfloat16 *base, int row_stride, int r0, int L,
                                              int tid) {
  const int r = r0 + tid;
  const uint4 *src = reinterpret_cast<const uint4 *>(base + (size_t)r * row_stride);
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<uint4 *>(dst + tile_off(128, tid, c * 8)) = r < L ? __ldg(src + c) : make_uint4(0, 0, 0, 0);
}
// all L rows (padded to NP) -> natural K-major tile [NP x 64] and / or transposed tile [64 x NP]
template <bool NAT, bool TR>
__device__ __forceinline__ void stage_all(uint8_t *nat, uint8_t *tr, const __nv_bfloat16 *base, int row_stride, int L,
                                          int NP, int tid) {
  for (int e = tid; e < (NP / 2) * 8; e += 128) {
    const int jp = e >> 3, c = e & 7, j0 = 2 * jp;
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
    if (j0 < L) a0 = __ldg(reinterpret_cast<const uint4 *>(base + (size_t)j0 * row_stride) + c);
    if (j0 + 1 < L) a1 = __ldg(reinterpret_cast<const uint4 *>(base + (size_t)(j0 + 1) * row_stride) + c);
    if (NAT) {
      *reinterpret_cast<uint4 *>(nat + tile_off(NP, j0, c * 8)) = a0;
      *reinterpret_cast<uint4 *>(nat + tile_off(NP, j0 + 1, c * 8)) = a1;
    }
    if (TR) {
      const unsigned short *e0 = reinterpret_cast<const unsigned short *>(&a0);
      const unsigned short *e1 = reinterpret_cast<const unsigned short *>(&a1);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<uint32_t *>(tr + tile_off(DH, c * 8 + i, j0)) = (uint32_t)e0[i] | ((uint32_t)e1[i] << 16);
    }
  }
}
__device__ __forceinline__ float gate_log(float wb, float w0, float w1, float w2, float w3, float w4, const float *l) {
  const float z = wb + w0 * l[0] + w1 * l[1] + w2 * l[2] + w3 * l[3] + w4 * l[4];
  return __logf(fmaxf(1.0f / (1.0f + __expf(-z)), 1e-6f));
}
constexpr float LOG_CLAMP = -13.815510557964274f;  // log(1e-6)

// ------------------------------------------------------------------------------------------------------------------
// B1: thread == query row
// ------------------------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(128, 1) attention_bwd_q_kernel(const BwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int NKP = NCH * 32;
  uint8_t *sQ = smem;                        // [128 x 64]
  uint8_t *sdO = sQ + 128 * DH * 2;          // [128 x 64]
  uint8_t *sK = sdO + 128 * DH * 2;          // [NKP x 64] rows = keys
  uint8_t *sV = sK + NKP * DH * 2;           // [NKP x 64]
  uint8_t *sKt = sV + NKP * DH * 2;          // [64 x NKP]
  uint8_t *sDS = sKt + DH * NKP * 2;         // [128 x NKP]
  uint64_t *mbar = reinterpret_cast<uint64_t *>(sDS + 128 * NKP * 2);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int E = a.H * DH;
  if (tid == 0) {
    mbar_init(mbar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  stage_all<true, true>(sK, sKt, a.k + (size_t)b * a.k_bs + h * DH, a.k_rs, a.Lk, NKP, tid);
  stage_all<true, false>(sV, nullptr, a.v + (size_t)b * a.v_bs + h * DH, a.v_rs, a.Lk, NKP, tid);
  uint32_t kmask[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int j = c * 32 + lane;
    kmask[c] = __ballot_sync(0xffffffffu, j < a.Lk && !(a.kpm && a.kpm[(size_t)b * a.Lk + j]));
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  constexpr uint32_t COL_S = 0, COL_DP = 160, COL_DQ = 320;
  uint32_t phase = 0;
  const bool gated = a.sw != nullptr;

  for (int q0 = 0; q0 < a.Lq; q0 += 128) {
    const int qi = q0 + tid;
    const bool qlive = qi < a.Lq;
    stage_tile128(sQ, a.q + (size_t)b * a.q_bs + h * DH, a.q_rs, q0, a.Lq, tid);
    stage_tile128(sdO, a.d_o + (size_t)b * a.Lq * E + h * DH, E, q0, a.Lq, tid);
    // D_i = sum_d dO_id * O_id
    float D = 0.f;
    if (qlive) {
      const uint4 *po = reinterpret_cast<const uint4 *>(a.o + ((size_t)b * a.Lq + qi) * E + h * DH);
      const uint4 *pd = reinterpret_cast<const uint4 *>(a.d_o + ((size_t)b * a.Lq + qi) * E + h * DH);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 x = __ldg(po + c), y = __ldg(pd + c);
        const __nv_bfloat162 *xa = reinterpret_cast<const __nv_bfloat162 *>(&x);
        const __nv_bfloat162 *ya = reinterpret_cast<const __nv_bfloat162 *>(&y);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 u = __bfloat1622float2(xa[i]), w = __bfloat1622float2(ya[i]);
          D += u.x * w.x + u.y * w.y;
        }
      }
      a.dvec[((size_t)b * a.H + h) * a.Lq + qi] = D;
    }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      const uint32_t idesc = make_idesc_bf16(128, NKP);
      const uint32_t aQ = smem_u32(sQ), aO = smem_u32(sdO), aK = smem_u32(sK), aV = smem_u32(sV);
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks)
        mma_bf16(tmem + COL_S, make_desc(aQ + ks * 4096, 2048, 128), make_desc(aK + ks * 2 * (NKP * 16), NKP * 16, 128),
                 idesc, ks > 0);
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks)
        mma_bf16(tmem + COL_DP, make_desc(aO + ks * 4096, 2048, 128), make_desc(aV + ks * 2 * (NKP * 16), NKP * 16, 128),
                 idesc, ks > 0);
      mma_commit(mbar);
    }
    float wb = 0.f, w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f, w4 = 0.f;
    const float *loc = nullptr;
    if (gated && qlive) {
      const float *w = a.sw + ((size_t)b * a.Lq + qi) * (a.H * 6) + h * 6;
      wb = w[0]; w0 = w[1]; w1 = w[2]; w2 = w[3]; w3 = w[4]; w4 = w[5];
      loc = a.locs + ((size_t)b * a.Lq + qi) * (size_t)a.Lk * 5;
    }
    const float lse = qlive ? a.lse[((size_t)b * a.H + h) * a.Lq + qi] : INFINITY;
    mbar_wait(mbar, phase);
    phase ^= 1u;
    fence_after_sync();
    float gb = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f, g4 = 0.f;
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      float s[32], dp[32];
      tmem_ld32(trow + COL_S + c * 32, s);
      tmem_ld32(trow + COL_DP + c * 32, dp);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int j = c * 32 + i;
        const bool on = ((kmask[c] >> i) & 1u) && qlive;
        float gl = 0.f;
        if (loc != nullptr && j < a.Lk) gl = gate_log(wb, w0, w1, w2, w3, w4, loc + (size_t)j * 5);
        const float p = on ? __expf(s[i] * a.scale + gl - lse) : 0.f;
        float dpi = dp[i];
        if (a.drop_thresh != 0u) {
          const unsigned long long idx = (((unsigned long long)b * a.H + h) * a.Lq + qi) * a.Lk + j;
          dpi = attn_rng::keep(a.seed, idx, a.drop_thresh) ? dpi * a.inv_keep : 0.f;
        }
        const float ds = p * (dpi - D);
        if (loc != nullptr && on && gl > LOG_CLAMP + 1e-3f) {  // clamp(sigmoid, 1e-6) inactive
          const float dz = ds * (1.0f - __expf(gl));  // d log(sigmoid(z)) / dz = 1 - sigmoid(z)
          const float *l = loc + (size_t)j * 5;
          gb += dz; g0 += dz * l[0]; g1 += dz * l[1]; g2 += dz * l[2]; g3 += dz * l[3]; g4 += dz * l[4];
        }
        s[i] = ds * a.scale;
      }
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        uint32_t w[4];
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) w[hh] = pack_bf16(s[qd * 8 + hh * 2], s[qd * 8 + hh * 2 + 1]);
        *reinterpret_cast<uint4 *>(sDS + tile_off(128, tid, c * 32 + qd * 8)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    if (a.dsw != nullptr && qlive) {
      float *g = a.dsw + ((size_t)b * a.Lq + qi) * (a.H * 6) + h * 6;
      g[0] = gb; g[1] = g0; g[2] = g1; g[3] = g2; g[4] = g3; g[5] = g4;
    }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      const uint32_t idesc = make_idesc_bf16(128, DH);
      const uint32_t aS = smem_u32(sDS), aT = smem_u32(sKt);
#pragma unroll
      for (int ks = 0; ks < NKP / 16; ++ks)
        mma_bf16(tmem + COL_DQ, make_desc(aS + ks * 4096, 2048, 128), make_desc(aT + ks * 2 * (DH * 16), DH * 16, 128), idesc,
                 ks > 0);
      mma_commit(mbar);
    }
    mbar_wait(mbar, phase);
    phase ^= 1u;
    fence_after_sync();
    {
      __nv_bfloat16 *o = a.dq + ((size_t)b * a.Lq + qi) * E + h * DH;
#pragma unroll
      for (int c0 = 0; c0 < DH; c0 += 32) {
        float v[32];
        tmem_ld32(trow + COL_DQ + c0, v);
        if (qlive) {
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            uint32_t w[4];
#pragma unroll
            for (int hh = 0; hh < 4; ++hh) w[hh] = pack_bf16(v[qd * 8 + hh * 2], v[qd * 8 + hh * 2 + 1]);
            *reinterpret_cast<uint4 *>(o + c0 + qd * 8) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
    }
    fence_before_sync();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------------------------
// B2: thread == key row
// ------------------------------------------------------------------------------------------------------------------
template <int NQCH>
__global__ void __launch_bounds__(128, 1) attention_bwd_kv_kernel(const BwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int NQP = NQCH * 32;
  uint8_t *sKr = smem;                        // [128 x 64]  rows = keys of this tile
  uint8_t *sVr = sKr + 128 * DH * 2;          // [128 x 64]
  uint8_t *sQn = sVr + 128 * DH * 2;          // [NQP x 64]  rows = queries
  uint8_t *sOn = sQn + NQP * DH * 2;          // [NQP x 64]  dO
  uint8_t *sQt = sOn + NQP * DH * 2;          // [64 x NQP]
  uint8_t *sOt = sQt + DH * NQP * 2;          // [64 x NQP]  dO^T
  uint8_t *sPT = sOt + DH * NQP * 2;          // [128 x NQP] P^T
  uint8_t *sDST = sPT + 128 * NQP * 2;        // [128 x NQP] scale * dS^T
  float *sLse = reinterpret_cast<float *>(sDST + 128 * NQP * 2);  // [NQP]
  float *sD = sLse + NQP;                                          // [NQP]
  float *sW = sD + NQP;                                            // [NQP][6]
  uint64_t *mbar = reinterpret_cast<uint64_t *>(sW + NQP * 6);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int E = a.H * DH;
  if (tid == 0) {
    mbar_init(mbar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  stage_all<true, true>(sQn, sQt, a.q + (size_t)b * a.q_bs + h * DH, a.q_rs, a.Lq, NQP, tid);
  stage_all<true, true>(sOn, sOt, a.d_o + (size_t)b * a.Lq * E + h * DH, E, a.Lq, NQP, tid);
  const bool gated = a.sw != nullptr;
  for (int i = tid; i < NQP; i += 128) {
    const bool live = i < a.Lq;
    sLse[i] = live ? a.lse[((size_t)b * a.H + h) * a.Lq + i] : INFINITY;
    sD[i] = live ? a.dvec[((size_t)b * a.H + h) * a.Lq + i] : 0.f;
    if (gated) {
      const float *w = a.sw + ((size_t)b * a.Lq + (live ? i : 0)) * (a.H * 6) + h * 6;
#pragma unroll
      for (int t = 0; t < 6; ++t) sW[i * 6 + t] = live ? w[t] : 0.f;
    }
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  constexpr uint32_t COL_ST = 0, COL_DPT = 160, COL_DV = 320, COL_DK = 384;
  uint32_t phase = 0;

  for (int k0 = 0; k0 < a.Lk; k0 += 128) {
    const int kj = k0 + tid;
    const bool klive = kj < a.Lk && !(a.kpm && a.kpm[(size_t)b * a.Lk + kj]);
    stage_tile128(sKr, a.k + (size_t)b * a.k_bs + h * DH, a.k_rs, k0, a.Lk, tid);
    stage_tile128(sVr, a.v + (size_t)b * a.v_bs + h * DH, a.v_rs, k0, a.Lk, tid);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      const uint32_t idesc = make_idesc_bf16(128, NQP);
      const uint32_t aK = smem_u32(sKr), aV = smem_u32(sVr), aQ = smem_u32(sQn), aO = smem_u32(sOn);
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks)
        mma_bf16(tmem + COL_ST, make_desc(aK + ks * 4096, 2048, 128), make_desc(aQ + ks * 2 * (NQP * 16), NQP * 16, 128),
                 idesc, ks > 0);
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks)
        mma_bf16(tmem + COL_DPT, make_desc(aV + ks * 4096, 2048, 128), make_desc(aO + ks * 2 * (NQP * 16), NQP * 16, 128),
                 idesc, ks > 0);
      mma_commit(mbar);
    }
    mbar_wait(mbar, phase);
    phase ^= 1u;
    fence_after_sync();
#pragma unroll 1
    for (int c = 0; c < NQCH; ++c) {
      float s[32], dp[32];
      tmem_ld32(trow + COL_ST + c * 32, s);
      tmem_ld32(trow + COL_DPT + c * 32, dp);
#pragma unroll
      for (int ii = 0; ii < 32; ++ii) {
        const int i = c * 32 + ii;  // query
        float p = 0.f;
        if (klive && i < a.Lq) {
          float gl = 0.f;
          if (gated) {
            const float *w = sW + i * 6;
            gl = gate_log(w[0], w[1], w[2], w[3], w[4], w[5], a.locs + (((size_t)b * a.Lq + i) * a.Lk + kj) * 5);
          }
          p = __expf(s[ii] * a.scale + gl - sLse[i]);
        }
        float pd = p, dpi = dp[ii];
        if (a.drop_thresh != 0u) {
          const unsigned long long idx = (((unsigned long long)b * a.H + h) * a.Lq + i) * a.Lk + kj;
          const float m = attn_rng::keep(a.seed, idx, a.drop_thresh) ? a.inv_keep : 0.f;
          pd = p * m;
          dpi *= m;
        }
        s[ii] = pd;
        dp[ii] = p * (dpi - sD[i]) * a.scale;
      }
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        uint32_t w[4], z[4];
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          w[hh] = pack_bf16(s[qd * 8 + hh * 2], s[qd * 8 + hh * 2 + 1]);
          z[hh] = pack_bf16(dp[qd * 8 + hh * 2], dp[qd * 8 + hh * 2 + 1]);
        }
        *reinterpret_cast<uint4 *>(sPT + tile_off(128, tid, c * 32 + qd * 8)) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4 *>(sDST + tile_off(128, tid, c * 32 + qd * 8)) = make_uint4(z[0], z[1], z[2], z[3]);
      }
    }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      const uint32_t idesc = make_idesc_bf16(128, DH);
      const uint32_t aP = smem_u32(sPT), aS = smem_u32(sDST), aOt = smem_u32(sOt), aQt = smem_u32(sQt);
#pragma unroll
      for (int ks = 0; ks < NQP / 16; ++ks)
        mma_bf16(tmem + COL_DV, make_desc(aP + ks * 4096, 2048, 128), make_desc(aOt + ks * 2 * (DH * 16), DH * 16, 128), idesc,
                 ks > 0);
#pragma unroll
      for (int ks = 0; ks < NQP / 16; ++ks)
        mma_bf16(tmem + COL_DK, make_desc(aS + ks * 4096, 2048, 128), make_desc(aQt + ks * 2 * (DH * 16), DH * 16, 128), idesc,
                 ks > 0);
      mma_commit(mbar);
    }
    mbar_wait(mbar, phase);
    phase ^= 1u;
    fence_after_sync();
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16 *o = (which ? a.dk : a.dv) + ((size_t)b * a.Lk + kj) * E + h * DH;
#pragma unroll
      for (int c0 = 0; c0 < DH; c0 += 32) {
        float v[32];
        tmem_ld32(trow + (which ? COL_DK : COL_DV) + c0, v);
        if (kj < a.Lk) {
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            uint32_t w[4];
#pragma unroll
            for (int hh = 0; hh < 4; ++hh) w[hh] = pack_bf16(v[qd * 8 + hh * 2], v[qd * 8 + hh * 2 + 1]);
            *reinterpret_cast<uint4 *>(o + c0 + qd * 8) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
    }
    fence_before_sync();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int NCH>
int launch_q(const BwdArgs &a, cudaStream_t st) {
  constexpr int NKP = NCH * 32;
  constexpr size_t smem = (size_t)2 * 128 * DH * 2 + (size_t)3 * NKP * DH * 2 + (size_t)128 * NKP * 2 + 32;
  auto kern = attention_bwd_q_kernel<NCH>;
  int rc = sv::cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (rc) return rc;
  kern<<<a.B * a.H, 128, smem, st>>>(a);
  return sv::after_launch();
}
template <int NQCH>
int launch_kv(const BwdArgs &a, cudaStream_t st) {
  constexpr int NQP = NQCH * 32;
  constexpr size_t smem = (size_t)2 * 128 * DH * 2 + (size_t)4 * NQP * DH * 2 + (size_t)2 * 128 * NQP * 2 + (size_t)NQP * 8 * 4 + 32;
  auto kern = attention_bwd_kv_kernel<NQCH>;
  int rc = sv::cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (rc) return rc;
  kern<<<a.B * a.H, 128, smem, st>>>(a);
  return sv::after_launch();
}

}  // namespace

extern "C" int sv_attention_bwd_bf16(const void *q, long long q_bs, int q_rs, const void *k, long long k_bs, int k_rs,
                                     const void *v, long long v_bs, int v_rs, const void *o, const void *d_o,
                                     const unsigned char *key_padding_mask, const float *spatial_w,
                                     const float *pairwise_locs, const float *lse, int B, int H, int Lq, int Lk,
                                     float scale, void *dq, void *dk, void *dv, float *d_spatial_w, float *dvec,
                                     void *stream) {
  return sv_attention_bwd_dropout_bf16(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, d_o, key_padding_mask, spatial_w,
                                       pairwise_locs, lse, B, H, Lq, Lk, scale, dq, dk, dv, d_spatial_w, dvec, 0.f, 0ull,
                                       stream);
}

extern "C" int sv_attention_bwd_dropout_bf16(const void *q, long long q_bs, int q_rs, const void *k, long long k_bs,
                                             int k_rs, const void *v, long long v_bs, int v_rs, const void *o,
                                             const void *d_o, const unsigned char *key_padding_mask,
                                             const float *spatial_w, const float *pairwise_locs, const float *lse, int B,
                                             int H, int Lq, int Lk, float scale, void *dq, void *dk, void *dv,
                                             float *d_spatial_w, float *dvec, float dropout_p, unsigned long long seed,
                                             void *stream) {
  if (dropout_p < 0.f || dropout_p >= 1.f) return SV_ERR_INVALID_ARG;
  if (B < 0 || H < 1 || Lq < 1 || Lk < 1 || Lq > 160 || Lk > 160) return SV_ERR_INVALID_ARG;
  if (B == 0) return SV_OK;
  if (!q || !k || !v || !o || !d_o || !lse || !dq || !dk || !dv || !dvec) return SV_ERR_INVALID_ARG;
  if ((q_rs % 8) || (k_rs % 8) || (v_rs % 8) || (q_bs % 8) || (k_bs % 8) || (v_bs % 8)) return SV_ERR_INVALID_ARG;
  if (spatial_w && (!pairwise_locs || !d_spatial_w)) return SV_ERR_INVALID_ARG;
  BwdArgs a;
  a.q = (const __nv_bfloat16 *)q; a.k = (const __nv_bfloat16 *)k; a.v = (const __nv_bfloat16 *)v;
  a.q_bs = q_bs; a.k_bs = k_bs; a.v_bs = v_bs; a.q_rs = q_rs; a.k_rs = k_rs; a.v_rs = v_rs;
  a.o = (const __nv_bfloat16 *)o; a.d_o = (const __nv_bfloat16 *)d_o;
  a.kpm = key_padding_mask; a.sw = spatial_w; a.locs = pairwise_locs; a.lse = lse;
  a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk; a.scale = scale;
  a.dq = (__nv_bfloat16 *)dq; a.dk = (__nv_bfloat16 *)dk; a.dv = (__nv_bfloat16 *)dv;
  a.dsw = d_spatial_w; a.dvec = dvec;
  a.drop_thresh = dropout_p > 0.f ? (unsigned)((double)dropout_p * 4294967296.0) : 0u;
  a.inv_keep = 1.0f / (1.0f - dropout_p);
  a.seed = seed;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  switch ((Lk + 31) / 32) {
    case 1: rc = launch_q<1>(a, st); break;
    case 2: rc = launch_q<2>(a, st); break;
    case 3: rc = launch_q<3>(a, st); break;
    case 4: rc = launch_q<4>(a, st); break;
    default: rc = launch_q<5>(a, st); break;
  }
  if (rc) return rc;
  switch ((Lq + 31) / 32) {
    case 1: return launch_kv<1>(a, st);
    case 2: return launch_kv<2>(a, st);
    case 3: return launch_kv<3>(a, st);
    case 4: return launch_kv<4>(a, st);
    default: return launch_kv<5>(a, st);
  }
}
