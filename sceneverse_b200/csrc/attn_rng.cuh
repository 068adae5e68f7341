// Counter-based dropout mask shared by the attention forward and backward kernels: the keep decision of attention weight
// (scene b, head h, query i, key j) depends only on (seed, linear index), so the backward regenerates exactly the mask
// the forward used (the reference draws it with nn.MultiheadAttention's dropout, transformers.py:22-24,118-120).
#pragma once
#include <stdint.h>

namespace attn_rng {
__device__ __forceinline__ bool keep(unsigned long long seed, unsigned long long idx, unsigned thresh) {
  unsigned long long x = idx + seed;  // murmur3 fmix64
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return (unsigned)x >= thresh;
}
}  // namespace attn_rng
