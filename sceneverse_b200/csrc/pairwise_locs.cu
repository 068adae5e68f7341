// calc_pairwise_locs, 'center' relation (reference: modules/utils.py:38-87): the 5-dim pairwise geometry
// [dist/max_dist, dz/dist, dist2d/dist, dy/dist2d, dx/dist2d] of the object centres of each scene, one fused
// kernel instead of ~12 elementwise ATen launches.  The max-distance normaliser is taken over ALL slots of the
// scene including padded (origin) objects, exactly like the reference (modules/utils.py:52-54).
#include "svcommon.h"
#include "svgps.h"

namespace {

__global__ void __launch_bounds__(256) pairwise_locs_kernel(const float *__restrict__ centers, int stride, int O, float eps,
                                                            int dist_norm, float *__restrict__ out) {
  extern __shared__ float sc[];  // [O][3]
  __shared__ float red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float *c = centers + (size_t)b * O * stride;
  for (int i = tid; i < O; i += 256) {
    sc[3 * i] = c[(size_t)i * stride];
    sc[3 * i + 1] = c[(size_t)i * stride + 1];
    sc[3 * i + 2] = c[(size_t)i * stride + 2];
  }
  __syncthreads();
  const int n = O * O;
  float mx = 0.f;
  if (dist_norm) {
    for (int e = tid; e < n; e += 256) {
      const int i = e / O, j = e - i * O;
      const float dx = sc[3 * i] - sc[3 * j], dy = sc[3 * i + 1] - sc[3 * j + 1], dz = sc[3 * i + 2] - sc[3 * j + 2];
      mx = fmaxf(mx, sqrtf(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)), eps)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  }
  float *o = out + (size_t)b * n * 5;
  for (int e = tid; e < n; e += 256) {
    const int i = e / O, j = e - i * O;
    const float dx = sc[3 * i] - sc[3 * j], dy = sc[3 * i + 1] - sc[3 * j + 1], dz = sc[3 * i + 2] - sc[3 * j + 2];
    const float xy = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    const float dist = sqrtf(__fadd_rn(__fadd_rn(xy, __fmul_rn(dz, dz)), eps));
    const float d2 = sqrtf(__fadd_rn(xy, eps));
    float *p = o + (size_t)e * 5;
    p[0] = dist_norm ? dist / mx : dist;
    p[1] = dz / dist;
    p[2] = d2 / dist;
    p[3] = dy / d2;
    p[4] = dx / d2;
  }
}

}  // namespace

extern "C" int sv_pairwise_locs_f32(const float *centers, int row_stride, int B, int O, float eps, int dist_norm,
                                    float *out, void *stream) {
  if (B < 0 || O < 0 || O > 1024 || row_stride < 3) return SV_ERR_INVALID_ARG;
  if (B == 0 || O == 0) return SV_OK;
  if (!centers || !out) return SV_ERR_INVALID_ARG;
  pairwise_locs_kernel<<<B, 256, (size_t)O * 3 * sizeof(float), (cudaStream_t)stream>>>(centers, row_stride, O, eps,
                                                                                        dist_norm, out);
  return sv::after_launch();
}
