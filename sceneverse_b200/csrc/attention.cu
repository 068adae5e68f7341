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
};

constexpr int DH = 64;

__global__ void __launch_bounds__(128, 2) attention_fwd_kernel(const AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int NKP = a.NKP;
  uint8_t *sQ = smem;                          // [128 x 64]   rows = queries, K = dh
  uint8_t *sK = sQ + 128 * DH * 2;             // [NKP x 64]   rows = keys,    K = dh
  uint8_t *sVt = sK + NKP * DH * 2;            // [64 x NKP]   rows = dh,      K = keys
  uint8_t *sP = sVt + DH * NKP * 2;            // [128 x NKP]  rows = queries, K = keys
  uint64_t *mbar = reinterpret_cast<uint64_t *>(sP + 128 * NKP * 2);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int q0 = blockIdx.y * 128;
  const int qi = q0 + tid;  // this thread's query row
  const bool qlive = qi < a.Lq;

  if (tid == 0) {
    mbar_init(mbar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);

  // ---- stage Q, K, V^T -----------------------------------------------------------------------------------------
  {
    uint4 row[8];
    if (qlive) {
      const uint4 *src = reinterpret_cast<const uint4 *>(a.q + (size_t)b * a.q_bs + (size_t)qi * a.q_rs + h * DH);
#pragma unroll
      for (int c = 0; c < 8; ++c) row[c] = __ldg(src + c);
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) row[c] = make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4 *>(sQ + tile_off(128, tid, c * 8)) = row[c];
  }
  for (int j = tid; j < NKP; j += 128) {
    uint4 kr[8], vr[8];
    if (j < a.Lk) {
      const uint4 *ks = reinterpret_cast<const uint4 *>(a.k + (size_t)b * a.k_bs + (size_t)j * a.k_rs + h * DH);
      const uint4 *vs = reinterpret_cast<const uint4 *>(a.v + (size_t)b * a.v_bs + (size_t)j * a.v_rs + h * DH);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        kr[c] = __ldg(ks + c);
        vr[c] = __ldg(vs + c);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) kr[c] = vr[c] = make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      *reinterpret_cast<uint4 *>(sK + tile_off(NKP, j, c * 8)) = kr[c];
      const unsigned short *e = reinterpret_cast<const unsigned short *>(&vr[c]);
#pragma unroll
      for (int i = 0; i < 8; ++i) *reinterpret_cast<unsigned short *>(sVt + tile_off(DH, c * 8 + i, j)) = e[i];
    }
  }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  constexpr uint32_t COL_S = 0, COL_O = 192;  // S: up to 160 columns, O: 64 columns

  // ---- S = Q K^T --------------------------------------------------------------------------------------------------
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, NKP);
    const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK);
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks)
      mma_bf16(tmem + COL_S, make_desc(aQ + ks * 2 * 2048, 2048, 128), make_desc(aK + ks * 2 * (NKP * 16), NKP * 16, 128),
               idesc, ks > 0);
    mma_commit(mbar);
  }
  // spatial gate parameters of this (query, head) while the MMA runs
  float wb = 0.f, w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f, w4 = 0.f;
  const float *loc = nullptr;
  if (a.sw != nullptr && qlive) {
    const float *w = a.sw + ((size_t)b * a.Lq + qi) * (a.SH * 6) + (a.SH == 1 ? 0 : h) * 6;
    wb = w[0]; w0 = w[1]; w1 = w[2]; w2 = w[3]; w3 = w[4]; w4 = w[5];
    loc = a.locs + ((size_t)b * a.Lq + qi) * (size_t)a.Lk * 5;
  }
  const unsigned char *kpm = a.kpm ? a.kpm + (size_t)b * a.Lk : nullptr;
  mbar_wait(mbar, 0);
  fence_after_sync();

  // ---- softmax over the keys of this thread's query row (two passes over TMEM: max, then exp/sum/store) ---------
  float mx = -INFINITY;
  for (int c0 = 0; c0 < NKP; c0 += 32) {
    float v[32];
    tmem_ld32(trow + COL_S + c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int j = c0 + i;
      float x = -INFINITY;
      if (j < a.Lk && !(kpm && kpm[j])) {
        x = v[i] * a.scale;
        if (loc) {
          const float *l = loc + (size_t)j * 5;
          const float z = wb + w0 * l[0] + w1 * l[1] + w2 * l[2] + w3 * l[3] + w4 * l[4];
          const float sg = 1.0f / (1.0f + __expf(-z));
          x += __logf(fmaxf(sg, 1e-6f));
        }
      }
      mx = fmaxf(mx, x);
    }
  }
  float sum = 0.f;
  for (int c0 = 0; c0 < NKP; c0 += 32) {
    float v[32];
    tmem_ld32(trow + COL_S + c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int j = c0 + i;
      float p = 0.f;
      if (j < a.Lk && !(kpm && kpm[j]) && mx > -INFINITY) {
        float x = v[i] * a.scale;
        if (loc) {
          const float *l = loc + (size_t)j * 5;
          const float z = wb + w0 * l[0] + w1 * l[1] + w2 * l[2] + w3 * l[3] + w4 * l[4];
          const float sg = 1.0f / (1.0f + __expf(-z));
          x += __logf(fmaxf(sg, 1e-6f));
        }
        p = __expf(x - mx);
      }
      sum += p;
      v[i] = p;
    }
    // unnormalised probabilities -> bf16 A operand; the 1/sum is applied to the fp32 output row
#pragma unroll
    for (int qd = 0; qd < 4; ++qd) {
      uint32_t w[4];
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) w[hh] = pack_bf16(v[qd * 8 + hh * 2], v[qd * 8 + hh * 2 + 1]);
      *reinterpret_cast<uint4 *>(sP + tile_off(128, tid, c0 + qd * 8)) = make_uint4(w[0], w[1], w[2], w[3]);
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
    for (int ks = 0; ks < NKP / 16; ++ks)
      mma_bf16(tmem + COL_O, make_desc(aP + ks * 2 * 2048, 2048, 128), make_desc(aV + ks * 2 * (DH * 16), DH * 16, 128),
               idesc, ks > 0);
    mma_commit(mbar);
  }
  mbar_wait(mbar, 1);
  fence_after_sync();
  {
    const float inv = sum > 0.f ? 1.0f / sum : 0.f;
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
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

}  // namespace

extern "C" int sv_attention_fwd_bf16(const void *q, long long q_bs, int q_rs, const void *k, long long k_bs, int k_rs,
                                     const void *v, long long v_bs, int v_rs, void *out, long long o_bs, int o_rs,
                                     const unsigned char *key_padding_mask, const float *spatial_w, int spatial_heads,
                                     const float *pairwise_locs, int B, int H, int Lq, int Lk, float scale,
                                     void *stream) {
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
  const size_t smem = (size_t)128 * DH * 2 + (size_t)a.NKP * DH * 2 * 2 + (size_t)128 * a.NKP * 2 + 32;
  int rc = sv::cuda_status(
      cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (rc) return rc;
  dim3 grid(B * H, (Lq + 127) / 128);
  attention_fwd_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(a);
  return sv::after_launch();
}
