// Fused attention forward on tcgen05 for the GPS shapes (sequence lengths <= 256 keys, head dim 64):
//   O = softmax( Q K^T / sqrt(64) + spatial_log_bias + key_padding_mask ) V          per (scene, head)
// Replaces the attention core of
//   * MultiHeadAttentionSpatial, 'cond' fusion (modules/layers/transformers.py:188-237): the per-(head, query)
//     spatial gate  log(clamp(sigmoid(w . pairwise_loc + b), 1e-6))  is computed on the fly from the 5-dim pairwise
//     geometry and the 6 language-conditioned weights — the (H,B,L,T) attention / loc_attn / mask tensors of the
//     reference are never materialised;
//   * nn.MultiheadAttention's scaled-dot-product core (joint self-attention of UnifiedSpatialCrossEncoderV2 and the
//     self/cross attention of the V1 / Entity decoders) in inference (no attention dropout).
// One CTA = one (scene, head, 128-query tile): Q, K, V^T and P live in shared memory in the canonical K-major UMMA
// layout, S = Q K^T and O = P V are tcgen05.mma with fp32 accumulators in TMEM, the softmax runs in registers with
// one thread per query row (== TMEM lane).  Rows may be strided (Q/K/V are usually slices of a packed projection).
#include <cuda_bf16.h>

#include "svcommon.h"
#include "svgps.h"
#include "attn_rng.cuh"
#include "tc05.cuh"

namespace {

using namespace tc05;

struct AttnArgs {
  const __nv_bfloat16 *q, *k, *v;  // (B,Lq,*), (B,Lk,*), (B,Lk,*): head h occupies columns [h*64, h*64+64)
  long long q_bs, k_bs, v_bs;      // batch strides (elements)
  int q_rs, k_rs, v_rs;            // row strides (elements)
  __nv_bfloat16 *out;              // (B,Lq,H*64)
  long long o_bs;
  int o_rs;
  const unsigned char *kpm;        // (B,Lk) 1 = masked key, or null
  const float *sw;                 // (B,Lq,SH*6) [bias, w1..w5] per spatial head, or null (plain attention)
  const float *locs;               // (B,Lq,Lk,5)
  int B, H, SH, Lq, Lk, NKP;       // NKP = Lk rounded up to a multiple of 32
  float scale;
  float *lse;                      // (B,H,Lq) log-sum-exp of the logits per query (for the backward), or null
  unsigned drop_thresh;            // dropout on the attention weights: keep iff hash >= thresh (0 = off)
  float inv_keep;                  // 1 / (1 - p)
  unsigned long long seed;
};

constexpr int DH = 64;

// gate(j) = log(clamp(sigmoid(w . loc_j + b), 1e-6)) for 4 consecutive keys from 5 float4 of the (Lk,5) row
struct Gate {
  float wb, w0, w1, w2, w3, w4;
  __device__ __forceinline__ float one(float l0, float l1, float l2, float l3, float l4) const {
    const float z = wb + w0 * l0 + w1 * l1 + w2 * l2 + w3 * l3 + w4 * l4;
    return __logf(fmaxf(1.0f / (1.0f + __expf(-z)), 1e-6f));
  }
};

template <int NCH>  // NCH = NKP / 32 (1..5)
__global__ void __launch_bounds__(128, 2) attention_fwd_kernel(const AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int NKP = NCH * 32;
  uint8_t *sQ = smem;                          // [128 x 64]   rows = queries, K = dh
  uint8_t *sK = sQ + 128 * DH * 2;             // [NKP x 64]   rows = keys,    K = dh
  uint8_t *sVt = sK + NKP * DH * 2;            // [64 x NKP]   rows = dh,      K = keys
  uint8_t *sP = sVt + DH * NKP * 2;            // [128 x NKP]  rows = queries, K = keys
  uint64_t *mbar = reinterpret_cast<uint64_t *>(sP + 128 * NKP * 2);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;

  if (tid == 0) {
    mbar_init(mbar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);

  // ---- stage K and V^T once per (scene, head) ---------------------------------------------------------------------
  // thread = (key pair, 8-column chunk): two keys' values of one column are packed into one 32-bit shared store
  for (int e = tid; e < (NKP / 2) * 8; e += 128) {
    const int jp = e >> 3, c = e & 7;
    const int j0 = 2 * jp;
    uint4 k0 = make_uint4(0, 0, 0, 0), k1 = k0, v0 = k0, v1 = k0;
    if (j0 < a.Lk) {
      k0 = __ldg(reinterpret_cast<const uint4 *>(a.k + (size_t)b * a.k_bs + (size_t)j0 * a.k_rs + h * DH) + c);
      v0 = __ldg(reinterpret_cast<const uint4 *>(a.v + (size_t)b * a.v_bs + (size_t)j0 * a.v_rs + h * DH) + c);
    }
    if (j0 + 1 < a.Lk) {
      k1 = __ldg(reinterpret_cast<const uint4 *>(a.k + (size_t)b * a.k_bs + (size_t)(j0 + 1) * a.k_rs + h * DH) + c);
      v1 = __ldg(reinterpret_cast<const uint4 *>(a.v + (size_t)b * a.v_bs + (size_t)(j0 + 1) * a.v_rs + h * DH) + c);
    }
    *reinterpret_cast<uint4 *>(sK + tile_off(NKP, j0, c * 8)) = k0;
    *reinterpret_cast<uint4 *>(sK + tile_off(NKP, j0 + 1, c * 8)) = k1;
    const unsigned short *e0 = reinterpret_cast<const unsigned short *>(&v0);
    const unsigned short *e1 = reinterpret_cast<const unsigned short *>(&v1);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<uint32_t *>(sVt + tile_off(DH, c * 8 + i, j0)) = (uint32_t)e0[i] | ((uint32_t)e1[i] << 16);
  }
  // key mask as one bit per key (bit set = key takes part), chunk c in kmask[c]
  uint32_t kmask[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int j = c * 32 + lane;
    const bool on = j < a.Lk && !(a.kpm && a.kpm[(size_t)b * a.Lk + j]);
    kmask[c] = __ballot_sync(0xffffffffu, on);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  constexpr uint32_t COL_S = 0, COL_O = 192;  // S: up to 160 columns, O: 64 columns
  uint32_t phase = 0;

  for (int q0 = 0; q0 < a.Lq; q0 += 128) {
    const int qi = q0 + tid;  // this thread's query row
    const bool qlive = qi < a.Lq;
    // ---- stage the Q tile ------------------------------------------------------------------------------------------
    {
      const uint4 *src = reinterpret_cast<const uint4 *>(a.q + (size_t)b * a.q_bs + (size_t)qi * a.q_rs + h * DH);
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4 *>(sQ + tile_off(128, tid, c * 8)) = qlive ? __ldg(src + c) : make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    // ---- S = Q K^T --------------------------------------------------------------------------------------------------
    if (tid == 0) {
      fence_after_sync();
      const uint32_t idesc = make_idesc_bf16(128, NKP);
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK);
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks)
        mma_bf16(tmem + COL_S, make_desc(aQ + ks * 4096, 2048, 128), make_desc(aK + ks * 2 * (NKP * 16), NKP * 16, 128),
                 idesc, ks > 0);
      mma_commit(mbar);
    }
    // spatial gate of this (query, head): computed ONCE per key while the MMA runs, kept in registers
    float gate[NKP];
    const bool gated = a.sw != nullptr;
    if (gated) {
#pragma unroll
      for (int j = 0; j < NKP; ++j) gate[j] = 0.f;
      if (qlive) {
        const float *w = a.sw + ((size_t)b * a.Lq + qi) * (a.SH * 6) + (a.SH == 1 ? 0 : h) * 6;
        const Gate g{w[0], w[1], w[2], w[3], w[4], w[5]};
        const float *loc = a.locs + ((size_t)b * a.Lq + qi) * (size_t)a.Lk * 5;
        if ((a.Lk & 3) == 0) {
          const float4 *l4 = reinterpret_cast<const float4 *>(loc);
#pragma unroll
          for (int j = 0; j < NKP; j += 4) {
            if (j < a.Lk) {
              const float4 A = __ldg(l4 + (j >> 2) * 5), B2 = __ldg(l4 + (j >> 2) * 5 + 1), C = __ldg(l4 + (j >> 2) * 5 + 2),
                           D = __ldg(l4 + (j >> 2) * 5 + 3), E = __ldg(l4 + (j >> 2) * 5 + 4);
              gate[j] = g.one(A.x, A.y, A.z, A.w, B2.x);
              gate[j + 1] = g.one(B2.y, B2.z, B2.w, C.x, C.y);
              gate[j + 2] = g.one(C.z, C.w, D.x, D.y, D.z);
              gate[j + 3] = g.one(D.w, E.x, E.y, E.z, E.w);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < NKP; ++j)
            if (j < a.Lk) gate[j] = g.one(loc[j * 5], loc[j * 5 + 1], loc[j * 5 + 2], loc[j * 5 + 3], loc[j * 5 + 4]);
        }
      }
    }
    mbar_wait(mbar, phase);
    phase ^= 1u;
    fence_after_sync();

    // ---- softmax over the keys of this thread's query row ----------------------------------------------------------
    // gated (<= 96 keys in practice): logits cached in `gate`; plain: second TMEM pass instead of 160 registers
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      float v[32];
      tmem_ld32(trow + COL_S + c * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const bool on = (kmask[c] >> i) & 1u;
        float x = v[i] * a.scale;
        if (gated) {
          x += gate[c * 32 + i];
          gate[c * 32 + i] = on ? x : -INFINITY;
        }
        mx = fmaxf(mx, on ? x : -INFINITY);
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      float v[32];
      if (!gated) tmem_ld32(trow + COL_S + c * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const bool on = (kmask[c] >> i) & 1u;
        const float x = gated ? gate[c * 32 + i] : v[i] * a.scale;
        const float p = (on && mx > -INFINITY) ? __expf(x - mx) : 0.f;
        sum += p;
        float pd = p;
        if (a.drop_thresh != 0u) {
          const unsigned long long idx = (((unsigned long long)b * a.H + h) * a.Lq + qi) * a.Lk + (c * 32 + i);
          pd = attn_rng::keep(a.seed, idx, a.drop_thresh) ? p * a.inv_keep : 0.f;
        }
        v[i] = pd;
      }
      // unnormalised probabilities -> bf16 A operand; 1/sum is applied to the fp32 output row
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        uint32_t w[4];
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) w[hh] = pack_bf16(v[qd * 8 + hh * 2], v[qd * 8 + hh * 2 + 1]);
        *reinterpret_cast<uint4 *>(sP + tile_off(128, tid, c * 32 + qd * 8)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();

    // ---- O = P V -----------------------------------------------------------------------------------------------------
    if (tid == 0) {
      fence_after_sync();
      const uint32_t idesc = make_idesc_bf16(128, DH);
      const uint32_t aP = smem_u32(sP), aV = smem_u32(sVt);
#pragma unroll
      for (int ks = 0; ks < NKP / 16; ++ks)
        mma_bf16(tmem + COL_O, make_desc(aP + ks * 4096, 2048, 128), make_desc(aV + ks * 2 * (DH * 16), DH * 16, 128),
                 idesc, ks > 0);
      mma_commit(mbar);
    }
    mbar_wait(mbar, phase);
    phase ^= 1u;
    fence_after_sync();
    {
      const float inv = sum > 0.f ? 1.0f / sum : 0.f;
      if (a.lse != nullptr && qlive) a.lse[((size_t)b * a.H + h) * a.Lq + qi] = sum > 0.f ? mx + __logf(sum) : INFINITY;
      __nv_bfloat16 *o = a.out + (size_t)b * a.o_bs + (size_t)qi * a.o_rs + h * DH;
#pragma unroll
      for (int c0 = 0; c0 < DH; c0 += 32) {
        float v[32];
        tmem_ld32(trow + COL_O + c0, v);
        if (qlive) {
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            uint32_t w[4];
#pragma unroll
            for (int hh = 0; hh < 4; ++hh) w[hh] = pack_bf16(v[qd * 8 + hh * 2] * inv, v[qd * 8 + hh * 2 + 1] * inv);
            *reinterpret_cast<uint4 *>(o + c0 + qd * 8) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
    }
    fence_before_sync();  // the next tile's MMA overwrites S / O
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

template <int NCH>
int launch_attn(const AttnArgs &a, cudaStream_t st) {
  constexpr int NKP = NCH * 32;
  constexpr size_t smem = (size_t)128 * DH * 2 + (size_t)NKP * DH * 2 * 2 + (size_t)128 * NKP * 2 + 32;
  auto kern = attention_fwd_kernel<NCH>;
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    int rc = sv::cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (rc) return rc;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    configured[dev] = true;
  }
  kern<<<a.B * a.H, 128, smem, st>>>(a);
  return sv::after_launch();
}

}  // namespace

extern "C" int sv_attention_fwd_bf16(const void *q, long long q_bs, int q_rs, const void *k, long long k_bs, int k_rs,
                                     const void *v, long long v_bs, int v_rs, void *out, long long o_bs, int o_rs,
                                     const unsigned char *key_padding_mask, const float *spatial_w, int spatial_heads,
                                     const float *pairwise_locs, int B, int H, int Lq, int Lk, float scale,
                                     void *stream) {
  return sv_attention_fwd_lse_bf16(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, out, o_bs, o_rs, key_padding_mask,
                                   spatial_w, spatial_heads, pairwise_locs, B, H, Lq, Lk, scale, nullptr, stream);
}

extern "C" int sv_attention_fwd_lse_bf16(const void *q, long long q_bs, int q_rs, const void *k, long long k_bs, int k_rs,
                                         const void *v, long long v_bs, int v_rs, void *out, long long o_bs, int o_rs,
                                         const unsigned char *key_padding_mask, const float *spatial_w,
                                         int spatial_heads, const float *pairwise_locs, int B, int H, int Lq, int Lk,
                                         float scale, float *lse, void *stream) {
  return sv_attention_fwd_dropout_bf16(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, out, o_bs, o_rs, key_padding_mask,
                                       spatial_w, spatial_heads, pairwise_locs, B, H, Lq, Lk, scale, lse, 0.f, 0ull, stream);
}

extern "C" int sv_attention_fwd_dropout_bf16(const void *q, long long q_bs, int q_rs, const void *k, long long k_bs,
                                             int k_rs, const void *v, long long v_bs, int v_rs, void *out, long long o_bs,
                                             int o_rs, const unsigned char *key_padding_mask, const float *spatial_w,
                                             int spatial_heads, const float *pairwise_locs, int B, int H, int Lq, int Lk,
                                             float scale, float *lse, float dropout_p, unsigned long long seed,
                                             void *stream) {
  if (dropout_p < 0.f || dropout_p >= 1.f) return SV_ERR_INVALID_ARG;
  if (B < 0 || H < 1 || Lq < 0 || Lk < 1 || Lk > 160) return SV_ERR_INVALID_ARG;
  if (B == 0 || Lq == 0) return SV_OK;
  if (!q || !k || !v || !out) return SV_ERR_INVALID_ARG;
  if ((q_rs % 8) || (k_rs % 8) || (v_rs % 8) || (o_rs % 8) || (q_bs % 8) || (k_bs % 8) || (v_bs % 8) || (o_bs % 8))
    return SV_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(k) & 15) ||
      (reinterpret_cast<uintptr_t>(v) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return SV_ERR_INVALID_ARG;
  if (spatial_w && (!pairwise_locs || (spatial_heads != 1 && spatial_heads != H))) return SV_ERR_INVALID_ARG;
  AttnArgs a;
  a.q = (const __nv_bfloat16 *)q; a.k = (const __nv_bfloat16 *)k; a.v = (const __nv_bfloat16 *)v;
  a.q_bs = q_bs; a.k_bs = k_bs; a.v_bs = v_bs; a.q_rs = q_rs; a.k_rs = k_rs; a.v_rs = v_rs;
  a.out = (__nv_bfloat16 *)out; a.o_bs = o_bs; a.o_rs = o_rs;
  a.kpm = key_padding_mask; a.sw = spatial_w; a.locs = pairwise_locs;
  a.B = B; a.H = H; a.SH = spatial_heads; a.Lq = Lq; a.Lk = Lk; a.NKP = (Lk + 31) / 32 * 32; a.scale = scale;
  a.lse = lse;
  a.drop_thresh = dropout_p > 0.f ? (unsigned)((double)dropout_p * 4294967296.0) : 0u;
  a.inv_keep = 1.0f / (1.0f - dropout_p);
  a.seed = seed;
  cudaStream_t st = (cudaStream_t)stream;
  switch (a.NKP / 32) {
    case 1: return launch_attn<1>(a, st);
    case 2: return launch_attn<2>(a, st);
    case 3: return launch_attn<3>(a, st);
    case 4: return launch_attn<4>(a, st);
    case 5: return launch_attn<5>(a, st);
    default: return SV_ERR_INVALID_ARG;
  }
}
