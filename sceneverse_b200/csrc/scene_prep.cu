// Device-side input pipeline of the GPS data path: what the reference does per object and per token in
// numpy / Python loops on the dataloader workers, as two kernels that build the model's `data_dict` tensors on the GPU
// from ragged raw scene data.
//   scene_prep_kernel  (reference: data/datasets/base.py:697-741 `_obj_processing_post` + dataset_wrapper.py:62-72 padding)
//     per object slot: obj_locs = [mean xyz, max - min] of the RAW points; subsample to P points — without replacement when
//     the object has >= P points, with replacement otherwise (np.random.choice(n, P, replace=n < P)); centre the SAMPLED
//     points on their mean, divide by their max norm (1 when < 1e-6); colours pass through; empty slots become all-ones
//     points, zero locs, mask 0.
//   token_mask_kernel  (reference: data/data_utils.py:76-121 `random_word`, `random_point_cloud`)
//     BERT masked-LM corruption of the caption (15 %: 80 % [MASK], 10 % random id, 10 % kept; label = original id, else -1)
//     and the object "semantic mask" coin flips.
// Randomness is a counter hash of (seed, slot / token, k): reproducible, order-independent, nothing stored.  Sampling
// without replacement is a keyed bijection of [0, 2^b) (xor / odd-multiply / xorshift rounds) with cycle walking onto [0, n):
// the first P images of 0..P-1 are P distinct uniform indices — no sort, no rejection table.
#include "svcommon.h"

namespace {

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}
__device__ __forceinline__ unsigned hash_u32(unsigned long long seed, unsigned long long a, unsigned long long b) {
  return (unsigned)(mix64(seed + a * 0x9E3779B97F4A7C15ULL + mix64(b + 0xD1B54A32D192ED03ULL)) >> 16);
}
// bijection of [0, 2^bits) keyed by (k1, k2, k3); bits >= 1
__device__ __forceinline__ unsigned perm_pow2(unsigned x, int bits, unsigned k1, unsigned k2, unsigned k3) {
  const unsigned mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
  const int sh = bits > 1 ? bits / 2 : 1;
  x = ((x ^ k1) * 0x9E3779B1u) & mask;
  x ^= x >> sh;
  x = ((x ^ k2) * 0x85EBCA6Bu) & mask;
  x ^= x >> sh;
  x = ((x ^ k3) * 0xC2B2AE35u) & mask;
  x ^= x >> sh;
  return x & mask;
}

struct PrepArgs {
  const float *raw;            // (total, 6) xyz rgb
  const long long *offsets;    // (S + 1) CSR over object slots; an empty range = padded slot
  int S, P;
  unsigned long long seed;
  float *fts;                  // (S, P, 6)
  float *locs;                 // (S, 6)
  unsigned char *masks;        // (S) 1 = real object
  int *sample_idx;             // (S, P) or null: the chosen raw indices (tests)
};

__device__ __forceinline__ double block_sum(double v, double *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  return t;
}
__device__ __forceinline__ float block_max(float v, float *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = -INFINITY;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t = fmaxf(t, red[w]);
  return t;
}

__global__ void __launch_bounds__(256) scene_prep_kernel(const PrepArgs a) {
  extern __shared__ __align__(16) float pts[];   // [P][6] sampled points
  __shared__ double redd[8];
  __shared__ float redf[8];
  const int s = blockIdx.x, tid = threadIdx.x;
  const long long beg = a.offsets[s], n = a.offsets[s + 1] - beg;
  float *out = a.fts + (size_t)s * a.P * 6;
  if (n <= 0) {   // padded slot: all-ones points (dataset_wrapper.py:62-64), zero locs, mask 0
    for (int i = tid; i < a.P * 6; i += 256) out[i] = 1.0f;
    if (tid < 6) a.locs[(size_t)s * 6 + tid] = 0.f;
    if (tid == 0) a.masks[s] = 0;
    if (a.sample_idx != nullptr)
      for (int i = tid; i < a.P; i += 256) a.sample_idx[(size_t)s * a.P + i] = -1;
    return;
  }
  const float *raw = a.raw + (size_t)beg * 6;
  // ---- obj_locs from ALL raw points: centre = mean, size = max - min --------------------------------------------------------
  double sx = 0, sy = 0, sz = 0;
  float mnx = INFINITY, mny = INFINITY, mnz = INFINITY, mxx = -INFINITY, mxy = -INFINITY, mxz = -INFINITY;
  for (long long i = tid; i < n; i += 256) {
    const float x = raw[i * 6], y = raw[i * 6 + 1], z = raw[i * 6 + 2];
    sx += x; sy += y; sz += z;
    mnx = fminf(mnx, x); mny = fminf(mny, y); mnz = fminf(mnz, z);
    mxx = fmaxf(mxx, x); mxy = fmaxf(mxy, y); mxz = fmaxf(mxz, z);
  }
  sx = block_sum(sx, redd); sy = block_sum(sy, redd); sz = block_sum(sz, redd);
  mxx = block_max(mxx, redf); mxy = block_max(mxy, redf); mxz = block_max(mxz, redf);
  mnx = -block_max(-mnx, redf); mny = -block_max(-mny, redf); mnz = -block_max(-mnz, redf);
  if (tid == 0) {
    float *l = a.locs + (size_t)s * 6;
    l[0] = (float)(sx / (double)n); l[1] = (float)(sy / (double)n); l[2] = (float)(sz / (double)n);
    l[3] = mxx - mnx; l[4] = mxy - mny; l[5] = mxz - mnz;
    a.masks[s] = 1;
  }
  // ---- subsample P points --------------------------------------------------------------------------------------------------
  int bits = 1;
  while ((1ll << bits) < n) ++bits;
  const unsigned k1 = hash_u32(a.seed, (unsigned long long)s, 1), k2 = hash_u32(a.seed, (unsigned long long)s, 2),
                 k3 = hash_u32(a.seed, (unsigned long long)s, 3);
  double cx = 0, cy = 0, cz = 0;
  for (int k = tid; k < a.P; k += 256) {
    long long idx;
    if (n < a.P) {   // with replacement
      idx = (long long)(((unsigned long long)hash_u32(a.seed, (unsigned long long)s, 16 + (unsigned long long)k) * (unsigned long long)n) >> 32);
    } else {         // without replacement: image of k under a keyed permutation of [0, n)
      unsigned x = perm_pow2((unsigned)k, bits, k1, k2, k3);
      while ((long long)x >= n) x = perm_pow2(x, bits, k1, k2, k3);
      idx = x;
    }
    if (a.sample_idx != nullptr) a.sample_idx[(size_t)s * a.P + k] = (int)idx;
    const float *p = raw + idx * 6;
#pragma unroll
    for (int c = 0; c < 6; ++c) pts[k * 6 + c] = p[c];
    cx += p[0]; cy += p[1]; cz += p[2];
  }
  cx = block_sum(cx, redd) / a.P; cy = block_sum(cy, redd) / a.P; cz = block_sum(cz, redd) / a.P;
  const float fx = (float)cx, fy = (float)cy, fz = (float)cz;
  // ---- centre on the sample mean, scale by the max norm ----------------------------------------------------------------------
  float md = 0.f;
  for (int k = tid; k < a.P; k += 256) {
    const float x = pts[k * 6] - fx, y = pts[k * 6 + 1] - fy, z = pts[k * 6 + 2] - fz;
    pts[k * 6] = x; pts[k * 6 + 1] = y; pts[k * 6 + 2] = z;
    md = fmaxf(md, sqrtf(x * x + y * y + z * z));
  }
  md = block_max(md, redf);
  const float inv = md < 1e-6f ? 1.0f : 1.0f / md;   // tiny clouds keep their coordinates (base.py:727-728)
  __syncthreads();
  for (int i = tid; i < a.P * 6; i += 256) {
    const int c = i % 6;
    out[i] = c < 3 ? pts[i] * inv : pts[i];
  }
}

struct MaskArgs {
  const long long *ids;            // (T) token ids
  const long long *attn;           // (T) attention mask 0/1
  long long *out_ids, *labels;     // (T)
  long long T;
  float ratio;
  long long mask_id, vocab;
  unsigned long long seed;
};
__global__ void __launch_bounds__(256) token_mask_kernel(const MaskArgs a) {
  const long long t = blockIdx.x * 256ll + threadIdx.x;
  if (t >= a.T) return;
  const long long tok = a.ids[t];
  long long out = tok, lab = -1;
  if (a.attn[t] != 0) {
    float u = (hash_u32(a.seed, (unsigned long long)t, 7) >> 8) * (1.0f / 16777216.0f);
    if (u < a.ratio) {
      u /= a.ratio;
      if (u < 0.8f) out = a.mask_id;
      else if (u < 0.9f) out = (long long)(((unsigned long long)hash_u32(a.seed, (unsigned long long)t, 11) * (unsigned long long)a.vocab) >> 32);
      lab = tok;
    }
  }
  a.out_ids[t] = out;
  a.labels[t] = lab;
}
__global__ void __launch_bounds__(256) coin_mask_kernel(const unsigned char *valid, unsigned char *out, long long n, float ratio,
                                                       unsigned long long seed) {
  const long long i = blockIdx.x * 256ll + threadIdx.x;
  if (i >= n) return;
  const float u = (hash_u32(seed, (unsigned long long)i, 13) >> 8) * (1.0f / 16777216.0f);
  out[i] = (valid[i] != 0 && !(u < ratio)) ? 1 : 0;
}

}  // namespace

extern "C" int sv_scene_prep_f32(const float *raw_points, const long long *slot_offsets, int n_slots, int P,
                                 unsigned long long seed, float *obj_fts, float *obj_locs, unsigned char *obj_masks,
                                 int *sample_idx, void *stream) {
  if (n_slots < 0 || P < 1 || P > 8192) return SV_ERR_INVALID_ARG;
  if (n_slots == 0) return SV_OK;
  if (!raw_points || !slot_offsets || !obj_fts || !obj_locs || !obj_masks) return SV_ERR_INVALID_ARG;
  const size_t smem = (size_t)P * 6 * sizeof(float);
  if (smem > 48 * 1024) {
    int rc = sv::cuda_status(cudaFuncSetAttribute(scene_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (rc) return rc;
  }
  PrepArgs a{raw_points, slot_offsets, n_slots, P, seed, obj_fts, obj_locs, obj_masks, sample_idx};
  scene_prep_kernel<<<n_slots, 256, smem, (cudaStream_t)stream>>>(a);
  return sv::after_launch();
}

extern "C" int sv_token_mask(const long long *ids, const long long *attention_mask, long long n_tokens, float mask_ratio,
                             long long mask_token_id, long long vocab_size, unsigned long long seed, long long *out_ids,
                             long long *labels, void *stream) {
  if (n_tokens < 0 || !(mask_ratio >= 0.f) || mask_ratio > 1.f || vocab_size < 1) return SV_ERR_INVALID_ARG;
  if (n_tokens == 0) return SV_OK;
  if (!ids || !attention_mask || !out_ids || !labels) return SV_ERR_INVALID_ARG;
  MaskArgs a{ids, attention_mask, out_ids, labels, n_tokens, mask_ratio, mask_token_id, vocab_size, seed};
  token_mask_kernel<<<(unsigned)((n_tokens + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  return sv::after_launch();
}

extern "C" int sv_coin_mask(const unsigned char *valid, long long n, float drop_ratio, unsigned long long seed,
                            unsigned char *out, void *stream) {
  if (n < 0 || !(drop_ratio >= 0.f) || drop_ratio > 1.f) return SV_ERR_INVALID_ARG;
  if (n == 0) return SV_OK;
  if (!valid || !out) return SV_ERR_INVALID_ARG;
  coin_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(valid, out, n, drop_ratio, seed);
  return sv::after_launch();
}
