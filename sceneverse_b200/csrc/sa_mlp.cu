// Fused PointNet++ set-abstraction MLP on the 5th-gen tensor cores (tcgen05 + TMEM), inference form
// (BatchNorm in eval mode folded to a per-channel affine, as in the frozen GPS backbone:
// pcd_openvocab_encoder.py:121-129 + all_pretrain.yaml `freeze: True`).
//
// Replaces, per set-abstraction level, the reference chain
//   QueryAndGroup (group_points x2, subtract centre, cat)  pointnet2_utils.py:345-356
//   SharedMLP = 3 x [Conv2d 1x1 -> BatchNorm2d -> ReLU]     pytorch_utils.py:11-36
//   max_pool2d over the nsample axis                       pointnet2_modules.py:65-71
// The grouped tensor (B, 3+C, npoint, nsample) is never materialised: rows are gathered straight
// into the UMMA operand layout in shared memory, the three GEMMs run back to back out of shared
// memory / TMEM, and the max over the neighbourhood is a running maximum in registers.
//
// Tiling: one accumulator tile = 128 rows = 128 different CENTRES for one fixed sample slot s; the
// CTA loops s = 0..nsample-1 over the same 128 centres ("super-tile"), so the neighbourhood max is
// a per-thread running max (thread == centre == TMEM lane) with no cross-lane traffic, and the
// BN scale is folded into the bf16 weights so that relu(acc + shift) is monotone in acc and the
// max can be taken on raw accumulators (affine + ReLU once per centre at the end).
// CTA = 256 threads = two warpgroups; both cover the 128 TMEM lanes and split the COLUMNS of every
// epilogue.  Thread 0 issues the MMAs; completion comes back through an mbarrier.
#include <cuda_bf16.h>

#include "svcommon.h"
#include "svgps.h"
#include "tc05.cuh"

namespace {

using namespace tc05;

template <int KF_, int K1P_, int N1_, int N2_, int N3_, int CPC_>
struct SaCfg {
  static constexpr int KF = KF_;    // feature channels gathered per neighbour (3 = rgb, 128 = SA1 output)
  static constexpr int K1P = K1P_;  // layer-1 K (3 + KF) padded to a multiple of 16
  static constexpr int N1 = N1_, N2 = N2_, N3 = N3_;
  static constexpr int CPC = CPC_;  // centres per cloud (npoint of this level)
  static constexpr int KMAX = K1P > N1 ? (K1P > N2 ? K1P : N2) : (N1 > N2 ? N1 : N2);
  static constexpr int A_BYTES = 128 * KMAX * 2;
  static constexpr int W1_BYTES = N1 * K1P * 2, W2_BYTES = N2 * N1 * 2, W3_BYTES = N3 * N2 * 2;
  static constexpr int SHIFT_BYTES = (N1 + N2 + N3) * 4;
  static constexpr int PARAM_BYTES = W1_BYTES + W2_BYTES + W3_BYTES + SHIFT_BYTES;
  static constexpr int TMEM_COLS = (N1 + N2 > N3 ? N1 + N2 : N3) <= 128 ? 128 : 256;
  static constexpr int SMEM_BYTES = A_BYTES + PARAM_BYTES + 64;
};

struct SaMlpArgs {
  const float *pts;       // level 1: (B,P,6) xyz+rgb f32 | level 2: (B,P,3) xyz f32 (P = points of this level's input)
  const void *feat;       // level 2: (B,P,KF) bf16 features of the input points (level 1: unused)
  const float *new_xyz;   // (B,CPC,3) centres
  const int *ball_idx;    // (B,CPC,NS)
  const void *params;     // packed [W1|W2|W3|shift1|shift2|shift3], canonical UMMA layout, BN scale folded
  void *out;              // (B,CPC,N3) bf16
  int B, P, NS;
};

// relu(acc + shift) for 32 accumulator columns -> 32 bf16 written as 4 x 16 B into the A tile (row r, cols c0..c0+31)
__device__ __forceinline__ void epilogue_to_smem(uint32_t taddr, const float *__restrict__ shift, uint8_t *sA, int r,
                                                 int c0) {
  float v[32];
  tmem_ld32(taddr, v);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t w[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int c = q * 8 + h * 2;
      const float a = fmaxf(v[c] + shift[c0 + c], 0.f), b = fmaxf(v[c + 1] + shift[c0 + c + 1], 0.f);
      w[h] = pack_bf16(a, b);
    }
    *reinterpret_cast<uint4 *>(sA + tile_off(128, r, c0 + q * 8)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

template <class Cfg, int LEVEL>
__global__ void __launch_bounds__(256, LEVEL == 1 ? 2 : 1) sa_mlp_kernel(const SaMlpArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *sA = smem;
  uint8_t *sW1 = smem + Cfg::A_BYTES;
  uint8_t *sW2 = sW1 + Cfg::W1_BYTES;
  uint8_t *sW3 = sW2 + Cfg::W2_BYTES;
  const float *sh1 = reinterpret_cast<const float *>(sW3 + Cfg::W3_BYTES);
  const float *sh2 = sh1 + Cfg::N1;
  const float *sh3 = sh2 + Cfg::N2;
  uint64_t *wbar = reinterpret_cast<uint64_t *>(smem + Cfg::A_BYTES + Cfg::PARAM_BYTES);
  uint64_t *mbar = wbar + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7;
  const int r = tid & 127;  // row of the tile == TMEM lane == centre within the super-tile

  if (tid == 0) {
    mbar_init(wbar, 1);
    mbar_init(mbar, 1);
    mbar_fence_init();
    mbar_expect_tx(wbar, Cfg::PARAM_BYTES);
    bulk_g2s(sW1, a.params, Cfg::PARAM_BYTES, wbar);
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  // zero the A tile once: padding columns of layer 1 stay zero for the whole kernel
  for (int e = tid; e < Cfg::K1P * 128 * 2 / 16; e += 256) reinterpret_cast<uint4 *>(sA)[e] = make_uint4(0, 0, 0, 0);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);  // this warp's 32 TMEM lanes
  mbar_wait(wbar, 0);

  constexpr uint32_t IDESC1 = make_idesc_bf16(128, Cfg::N1), IDESC2 = make_idesc_bf16(128, Cfg::N2),
                     IDESC3 = make_idesc_bf16(128, Cfg::N3);
  constexpr uint32_t COL1 = 0, COL2 = Cfg::N1, COL3 = 0;  // layer 3 reuses the columns of layers 1-2
  const uint32_t aA = smem_u32(sA), aW1 = smem_u32(sW1), aW2 = smem_u32(sW2), aW3 = smem_u32(sW3);
  uint32_t phase = 0;

  const int n_centres = a.B * Cfg::CPC;
  const int n_super = (n_centres + 127) / 128;
  for (int st = blockIdx.x; st < n_super; st += gridDim.x) {
    const int cg = st * 128 + r;  // global centre index
    const bool live = cg < n_centres;
    const int cloud = live ? cg / Cfg::CPC : 0;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (live) {
      const float *c = a.new_xyz + (size_t)cg * 3;
      cx = c[0]; cy = c[1]; cz = c[2];
    }
    const int *my_idx = a.ball_idx + (size_t)cg * a.NS;
    float runmax[Cfg::N3 / 2];
#pragma unroll
    for (int i = 0; i < Cfg::N3 / 2; ++i) runmax[i] = -INFINITY;

    for (int s = 0; s < a.NS; ++s) {
      // ---- gather row r = (centre cg, sample s) into the A tile -------------------------------------
      if (LEVEL == 1) {
        if (wg == 0) {
          uint4 row = make_uint4(0, 0, 0, 0);
          if (live) {
            const int k = __ldg(my_idx + s);
            const float2 *p = reinterpret_cast<const float2 *>(a.pts + ((size_t)cloud * a.P + k) * 6);
            const float2 p0 = __ldg(p), p1 = __ldg(p + 1), p2 = __ldg(p + 2);  // x y | z r | g b
            row.x = pack_bf16(p0.x - cx, p0.y - cy);
            row.y = pack_bf16(p1.x - cz, p1.y);
            row.z = pack_bf16(p2.x, p2.y);
          }
          *reinterpret_cast<uint4 *>(sA + tile_off(128, r, 0)) = row;
          // columns 8..15 are K padding; the layer-1/2 epilogues reuse this space, so re-zero it every tile
          *reinterpret_cast<uint4 *>(sA + tile_off(128, r, 8)) = make_uint4(0, 0, 0, 0);
        }
      } else {
        // columns [0,KF) = features of the neighbour, [KF,KF+3) = xyz - centre (weights permuted to match)
        int k = 0;
        if (live) k = __ldg(my_idx + s);
        constexpr int CH = Cfg::KF / 8;  // 16-byte chunks of the feature row
        const uint4 *f = reinterpret_cast<const uint4 *>(a.feat) + ((size_t)cloud * a.P + k) * CH;
#pragma unroll
        for (int q = wg * (CH / 2); q < (wg + 1) * (CH / 2); ++q) {
          const uint4 v = live ? __ldg(f + q) : make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4 *>(sA + tile_off(128, r, q * 8)) = v;
        }
        if (wg == 1) {
          uint4 row = make_uint4(0, 0, 0, 0);
          if (live) {
            const float *p = a.pts + ((size_t)cloud * a.P + k) * 3;
            row.x = pack_bf16(__ldg(p) - cx, __ldg(p + 1) - cy);
            row.y = pack_bf16(__ldg(p + 2) - cz, 0.f);
          }
          *reinterpret_cast<uint4 *>(sA + tile_off(128, r, Cfg::KF)) = row;
        }
      }
      fence_proxy_async_smem();
      fence_before_sync();
      __syncthreads();
      // ---- layer 1 ------------------------------------------------------------------------------------------
      if (tid == 0) {
        fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < Cfg::K1P / 16; ++ks)
          mma_bf16(tmem + COL1, make_desc(aA + ks * 2 * 2048, 2048, 128),
                   make_desc(aW1 + ks * 2 * (Cfg::N1 * 16), Cfg::N1 * 16, 128), IDESC1, ks > 0);
        mma_commit(mbar);
      }
      mbar_wait(mbar, phase);
      phase ^= 1;
      fence_after_sync();
#pragma unroll
      for (int c0 = wg * (Cfg::N1 / 2); c0 < (wg + 1) * (Cfg::N1 / 2); c0 += 32)
        epilogue_to_smem(trow + COL1 + c0, sh1, sA, r, c0);
      fence_proxy_async_smem();
      fence_before_sync();
      __syncthreads();
      // ---- layer 2 ------------------------------------------------------------------------------------------
      if (tid == 0) {
        fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < Cfg::N1 / 16; ++ks)
          mma_bf16(tmem + COL2, make_desc(aA + ks * 2 * 2048, 2048, 128),
                   make_desc(aW2 + ks * 2 * (Cfg::N2 * 16), Cfg::N2 * 16, 128), IDESC2, ks > 0);
        mma_commit(mbar);
      }
      mbar_wait(mbar, phase);
      phase ^= 1;
      fence_after_sync();
#pragma unroll
      for (int c0 = wg * (Cfg::N2 / 2); c0 < (wg + 1) * (Cfg::N2 / 2); c0 += 32)
        epilogue_to_smem(trow + COL2 + c0, sh2, sA, r, c0);
      fence_proxy_async_smem();
      fence_before_sync();
      __syncthreads();
      // ---- layer 3 + running max over the neighbourhood -----------------------------------------------------
      if (tid == 0) {
        fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < Cfg::N2 / 16; ++ks)
          mma_bf16(tmem + COL3, make_desc(aA + ks * 2 * 2048, 2048, 128),
                   make_desc(aW3 + ks * 2 * (Cfg::N3 * 16), Cfg::N3 * 16, 128), IDESC3, ks > 0);
        mma_commit(mbar);
      }
      mbar_wait(mbar, phase);
      phase ^= 1;
      fence_after_sync();
#pragma unroll
      for (int cc = 0; cc < Cfg::N3 / 64; ++cc) {
        float v[32];
        tmem_ld32(trow + COL3 + wg * (Cfg::N3 / 2) + cc * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) runmax[cc * 32 + i] = fmaxf(runmax[cc * 32 + i], v[i]);
      }
      // (the next gather overwrites sA / the next MMA overwrites TMEM only after the __syncthreads that follows it)
      fence_before_sync();
    }
    // ---- finalize: relu(max + shift3) -> bf16 row of this centre -------------------------------------------
    if (live) {
      uint4 *o = reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(a.out) + (size_t)cg * Cfg::N3 +
                                            wg * (Cfg::N3 / 2));
#pragma unroll
      for (int q = 0; q < Cfg::N3 / 16; ++q) {
        uint32_t w[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int c = q * 8 + h * 2;
          const int gc = wg * (Cfg::N3 / 2) + c;
          w[h] = pack_bf16(fmaxf(runmax[c] + sh3[gc], 0.f), fmaxf(runmax[c + 1] + sh3[gc + 1], 0.f));
        }
        o[q] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

using Sa1 = SaCfg<3, 16, 64, 64, 128, 32>;
using Sa2 = SaCfg<128, 144, 128, 128, 256, 16>;

template <class Cfg, int LEVEL>
int launch_sa(const SaMlpArgs &a, cudaStream_t st) {
  auto kern = sa_mlp_kernel<Cfg, LEVEL>;
  int rc = sv::cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  if (rc) return rc;
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  rc = sv::cuda_status(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, Cfg::SMEM_BYTES));
  if (rc) return rc;
  if (per_sm < 1) per_sm = 1;
  if (per_sm * Cfg::TMEM_COLS > 512) per_sm = 512 / Cfg::TMEM_COLS;  // TMEM columns are not part of the occupancy query
  const int n_super = (a.B * Cfg::CPC + 127) / 128;
  int grid = sms * per_sm;
  if (grid > n_super) grid = n_super;
  kern<<<grid, 256, Cfg::SMEM_BYTES, st>>>(a);
  return sv::after_launch();
}

}  // namespace

extern "C" {

int sv_sa_mlp_param_bytes(int level) { return level == 1 ? Sa1::PARAM_BYTES : level == 2 ? Sa2::PARAM_BYTES : -1; }

int sv_sa1_mlp_bf16(const float *pts, const float *new_xyz, const int *ball_idx, const void *params, int B, int P,
                    int nsample, void *out_feat, void *stream) {
  if (B < 0 || P < 1 || nsample < 1) return SV_ERR_INVALID_ARG;
  if (B == 0) return SV_OK;
  if (!pts || !new_xyz || !ball_idx || !params || !out_feat) return SV_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(params) & 15) || (reinterpret_cast<uintptr_t>(out_feat) & 15) ||
      (reinterpret_cast<uintptr_t>(pts) & 7))
    return SV_ERR_INVALID_ARG;
  SaMlpArgs a{pts, nullptr, new_xyz, ball_idx, params, out_feat, B, P, nsample};
  return launch_sa<Sa1, 1>(a, (cudaStream_t)stream);
}

int sv_sa2_mlp_bf16(const float *xyz, const void *feat, const float *new_xyz, const int *ball_idx, const void *params,
                    int B, int P, int nsample, void *out_feat, void *stream) {
  if (B < 0 || P < 1 || nsample < 1) return SV_ERR_INVALID_ARG;
  if (B == 0) return SV_OK;
  if (!xyz || !feat || !new_xyz || !ball_idx || !params || !out_feat) return SV_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(params) & 15) || (reinterpret_cast<uintptr_t>(out_feat) & 15) ||
      (reinterpret_cast<uintptr_t>(feat) & 15))
    return SV_ERR_INVALID_ARG;
  SaMlpArgs a{xyz, feat, new_xyz, ball_idx, params, out_feat, B, P, nsample};
  return launch_sa<Sa2, 2>(a, (cudaStream_t)stream);
}

}  // extern "C"
