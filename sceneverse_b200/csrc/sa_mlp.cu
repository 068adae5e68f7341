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
// CTA = NCG warpgroups of 128 threads (SA1: 2, two CTAs per SM; SA2: 4, one CTA per SM): every warpgroup covers the 128 TMEM lanes (rows) and owns 1/NCG of the
// COLUMNS of every epilogue, so 16 warps per CTA hide the TMEM-load / shared-memory latencies of the short epilogue
// phases.  One EXTRA warp is the MMA issuer: the epilogue threads never meet in a CTA-wide barrier — each arrives on the
// tile's `a_ready` mbarrier when its part of the next A operand is in shared memory and moves straight on to the other
// tile in flight; the issuer waits for `a_ready`, issues the layer's MMAs and commits to `mma_done`.
#include <cuda_bf16.h>

#include "svcommon.h"
#include "svgps.h"
#include "tc05.cuh"

namespace {

using namespace tc05;

template <int KF_, int K1P_, int N1_, int N2_, int N3_, int CPC_, int NCG_>
struct SaCfg {
  static constexpr int NCG = NCG_;  // column groups (warpgroups of 128 threads) per CTA
  static constexpr int NTHREADS = 128 * NCG_;         // epilogue / gather threads
  // SA2 (tensor-pipe bound, one CTA per SM) gets a dedicated MMA issuer warp; SA1 (CUDA-core bound, two CTAs per SM,
  // measured slower with the extra warp's register cost) keeps a CTA barrier + thread 0 as the issuer
  static constexpr bool ISSUER = NCG_ == 4;
  static constexpr int NTHREADS_ALL = 128 * NCG_ + (ISSUER ? 32 : 0);
  static constexpr int KF = KF_;    // feature channels gathered per neighbour (3 = rgb, 128 = SA1 output)
  static constexpr int K1P = K1P_;  // layer-1 K (3 + KF) padded to a multiple of 16
  static constexpr int N1 = N1_, N2 = N2_, N3 = N3_;
  static constexpr int CPC = CPC_;  // centres per cloud (npoint of this level)
  // Operand layout: columns [0, KSW) of a layer's K live in SWIZZLE_128B slabs of 64 columns, the remaining (< 64)
  // columns — only the layer-1 tail (SA1: all 16, SA2: xyz, columns 128..143) — in a no-swizzle tail block behind them.
  static constexpr int KSW1 = (K1P / 64) * 64, KT1 = K1P - KSW1;  // layer 1: swizzled / tail columns
  static constexpr int SWMAX = (N1 > N2 ? N1 : N2) > KSW1 ? (N1 > N2 ? N1 : N2) : KSW1;
  static constexpr int A_TAIL_OFF = (SWMAX / 64) * (128 * 128);   // tail block of the A tile
  static constexpr int A_BYTES = A_TAIL_OFF + 128 * 16 * 2;        // + 16 tail columns
  static constexpr int W1_BYTES = N1 * K1P * 2, W2_BYTES = N2 * N1 * 2, W3_BYTES = N3 * N2 * 2;
  static constexpr int SHIFT_BYTES = (N1 + N2 + N3) * 4;
  static constexpr int PARAM_BYTES = W1_BYTES + W2_BYTES + W3_BYTES + SHIFT_BYTES;
  static constexpr int TMEM_COLS = (N1 + N2 > N3 ? N1 + N2 : N3) <= 128 ? 128 : 256;
  static constexpr int SMEM_BYTES = 2 * A_BYTES + PARAM_BYTES + 64;  // two tiles in flight
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


// 32 lanes x 16 columns TMEM load
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// relu(acc + shift) for NC (16 or 32) accumulator columns -> bf16, 16-byte chunks into the A tile (row r, cols c0..)
template <int NC>
__device__ __forceinline__ void epilogue_to_smem(uint32_t taddr, const float *__restrict__ shift, uint8_t *sA, int r,
                                                 int c0) {
  float v[NC];
  if constexpr (NC == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
#pragma unroll
  for (int q = 0; q < NC / 8; ++q) {
    uint32_t w[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int c = q * 8 + h * 2;
      const float a = fmaxf(v[c] + shift[c0 + c], 0.f), b = fmaxf(v[c + 1] + shift[c0 + c + 1], 0.f);
      w[h] = pack_bf16(a, b);
    }
    *reinterpret_cast<uint4 *>(sA + tile_off_sw128(128, r, c0 + q * 8)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// 16-byte async copy global -> shared (LDGSTS); src_bytes = 0 zero-fills
__device__ __forceinline__ void cp_async16(void *dst, const void *src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Two tiles (sample slots s and s+1 of the same 128 centres) are in flight per CTA, each with its own
// A buffer, TMEM column range and mbarrier, so the tensor pipe works on one while the CUDA cores run the
// epilogue / gather of the other.
template <class Cfg, int LEVEL>
__global__ void __launch_bounds__(Cfg::NTHREADS_ALL, Cfg::NCG == 2 ? 2 : 1) sa_mlp_kernel(const SaMlpArgs a) {
  constexpr int NCG = Cfg::NCG, NTHREADS = Cfg::NTHREADS;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *sW1 = smem + 2 * Cfg::A_BYTES;
  uint8_t *sW2 = sW1 + Cfg::W1_BYTES;
  uint8_t *sW3 = sW2 + Cfg::W2_BYTES;
  const float *sh1 = reinterpret_cast<const float *>(sW3 + Cfg::W3_BYTES);
  const float *sh2 = sh1 + Cfg::N1;
  const float *sh3 = sh2 + Cfg::N2;
  uint64_t *wbar = reinterpret_cast<uint64_t *>(smem + 2 * Cfg::A_BYTES + Cfg::PARAM_BYTES);
  uint64_t *mbar = wbar + 1;    // [2] mma_done: the issuer's tcgen05.commit of the tile in buffer b
  uint64_t *abar = mbar + 2;    // [2] a_ready: every epilogue thread has written its part of buffer b's next A operand
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(abar + 2);

  const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7;  // wg = column group of this thread
  const int r = tid & 127;  // row of the tile == TMEM lane == centre within the super-tile
  const bool is_issuer = Cfg::ISSUER && warp == 4 * NCG;  // the extra warp
  const bool tmem_owner = Cfg::ISSUER ? is_issuer : warp == 1;

  if (tid == 0) {
    mbar_init(wbar, 1);
    mbar_init(mbar, 1);
    mbar_init(mbar + 1, 1);
    mbar_init(abar, NTHREADS);
    mbar_init(abar + 1, NTHREADS);
    mbar_fence_init();
    mbar_expect_tx(wbar, Cfg::PARAM_BYTES);
    bulk_g2s(sW1, a.params, Cfg::PARAM_BYTES, wbar);
  }
  if (tmem_owner) tmem_alloc<2 * Cfg::TMEM_COLS>(tmem_slot);
  // zero both A tiles once: K-padding columns that no gather / epilogue writes stay zero for the whole kernel
  for (int e = tid; e < 2 * Cfg::A_BYTES / 16; e += Cfg::NTHREADS_ALL) reinterpret_cast<uint4 *>(smem)[e] = make_uint4(0, 0, 0, 0);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;  // this warp's 32 TMEM lanes
  mbar_wait(wbar, 0);

  constexpr uint32_t IDESC1 = make_idesc_bf16(128, Cfg::N1), IDESC2 = make_idesc_bf16(128, Cfg::N2),
                     IDESC3 = make_idesc_bf16(128, Cfg::N3);
  const uint32_t aW1 = smem_u32(sW1), aW2 = smem_u32(sW2), aW3 = smem_u32(sW3);
  uint32_t phase[2] = {0u, 0u};

  // issue layer L (1..3) of buffer b, after a_ready[b]: called by a whole converged warp, elect.sync picks the issuing lane
  // (tc05.cuh: warp-converged issue)
  auto issue = [&](int L, int b) {
    fence_after_sync();
    const uint32_t aA = smem_u32(smem + b * Cfg::A_BYTES);
    const uint32_t tb = tmem + b * Cfg::TMEM_COLS;
    // K-step ks of a layer with K columns: swizzled slab addressing for columns < KSW, no-swizzle tail behind
    auto step = [&](uint32_t acc, uint32_t aW, int N, int KSW, int ks, uint32_t idesc) {
      const int c0 = ks * 16;
      if (c0 < KSW) {
        mma_bf16_e(acc, make_desc_sw128(aA + (c0 >> 6) * (128 * 128) + ((c0 >> 4) & 3) * 32),
                 make_desc_sw128(aW + (c0 >> 6) * (N * 128) + ((c0 >> 4) & 3) * 32), idesc, ks > 0);
      } else {
        const int kt = (c0 - KSW) >> 4;
        mma_bf16_e(acc, make_desc(aA + Cfg::A_TAIL_OFF + kt * 4096, 2048, 128),
                 make_desc(aW + (KSW >> 6) * (N * 128) + kt * 2 * (N * 16), N * 16, 128), idesc, ks > 0);
      }
    };
    if (L == 1) {
#pragma unroll
      for (int ks = 0; ks < Cfg::K1P / 16; ++ks) step(tb, aW1, Cfg::N1, Cfg::KSW1, ks, IDESC1);
    } else if (L == 2) {
#pragma unroll
      for (int ks = 0; ks < Cfg::N1 / 16; ++ks) step(tb + Cfg::N1, aW2, Cfg::N2, Cfg::N1, ks, IDESC2);
    } else {
#pragma unroll
      for (int ks = 0; ks < Cfg::N2 / 16; ++ks) step(tb, aW3, Cfg::N3, Cfg::N2, ks, IDESC3);
    }
    mma_commit_e(mbar + b);
  };
  auto wait_mma = [&](int b) {
    mbar_wait(mbar + b, phase[b]);
    phase[b] ^= 1u;
    fence_after_sync();
  };
  // epilogue thread: its smem writes (generic proxy) and TMEM reads of buffer b are done -> arrive on a_ready[b]
  // (without the issuer warp: CTA barrier, then thread 0 issues layer L itself)
  auto publish = [&](int L, int b) {
    fence_proxy_async_smem();
    fence_before_sync();
    if constexpr (Cfg::ISSUER) {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(abar + b)) : "memory");
    } else {
      __syncthreads();
      if (tid < 32) issue(L, b);
    }
  };
  uint32_t aphase[2] = {0u, 0u};
  // issuer lane: wait until every epilogue thread arrived for buffer b, then issue layer L
  auto wait_and_issue = [&](int L, int b) {
    mbar_wait(abar + b, aphase[b]);
    aphase[b] ^= 1u;
    issue(L, b);
  };

  const int n_centres = a.B * Cfg::CPC;
  const int n_super = (n_centres + 127) / 128;
  const int NS = a.NS;
  for (int st = blockIdx.x; st < n_super; st += gridDim.x) {
    if (is_issuer) {
      // same event order as the epilogue threads' program order: (L1,b0) (L1,b1) | per pair: (L2,b0) (L2,b1) (L3,b0)
      // (L3,b1) (L1',b0) (L1',b1)
      {
        wait_and_issue(1, 0);
        if (1 < NS) wait_and_issue(1, 1);
        for (int s = 0; s < NS; s += 2) {
          const bool has1 = s + 1 < NS;
#pragma unroll
          for (int L = 1; L <= 2; ++L) {
            wait_and_issue(L + 1, 0);
            if (has1) wait_and_issue(L + 1, 1);
          }
          if (s + 2 < NS) wait_and_issue(1, 0);
          if (has1 && s + 3 < NS) wait_and_issue(1, 1);
        }
      }
      continue;
    }
    const int cg = st * 128 + r;  // global centre index
    const bool live = cg < n_centres;
    const int cloud = live ? cg / Cfg::CPC : 0;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (live) {
      const float *c = a.new_xyz + (size_t)cg * 3;
      cx = c[0]; cy = c[1]; cz = c[2];
    }
    const int *my_idx = a.ball_idx + (size_t)cg * NS;
    // running neighbourhood maximum of the raw layer-3 accumulators, kept as packed bf16 pairs: rounding is monotone, so
    // max(round(x)) == round(max(x)) — half the registers of an fp32 running max (they pay for the issuer warp)
    // (only where the issuer warp squeezes the register budget; SA1 keeps an fp32 running max)
    constexpr bool PACKED = Cfg::ISSUER;
    __nv_bfloat162 runmax[PACKED ? Cfg::N3 / NCG / 2 : 1];
    float runmaxf[PACKED ? 1 : Cfg::N3 / NCG];
    if constexpr (PACKED) {
#pragma unroll
      for (int i = 0; i < Cfg::N3 / NCG / 2; ++i) runmax[i] = __floats2bfloat162_rn(-INFINITY, -INFINITY);
    } else {
#pragma unroll
      for (int i = 0; i < Cfg::N3 / NCG; ++i) runmaxf[i] = -INFINITY;
    }

    // ---- gather of row r = (centre cg, sample s) into A buffer b, split in "load" (early) and "store" -----
    struct Pre { int k; float v0, v1, v2, v3, v4, v5; };
    auto preload = [&](int s) {
      Pre p{0, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (live && s < NS) {
        p.k = __ldg(my_idx + s);
        if (LEVEL == 1) {
          const float2 *q = reinterpret_cast<const float2 *>(a.pts + ((size_t)cloud * a.P + p.k) * 6);
          const float2 p0 = __ldg(q), p1 = __ldg(q + 1), p2 = __ldg(q + 2);  // x y | z r | g b
          p.v0 = p0.x; p.v1 = p0.y; p.v2 = p1.x; p.v3 = p1.y; p.v4 = p2.x; p.v5 = p2.y;
        } else if (wg == NCG - 1) {
          const float *q = a.pts + ((size_t)cloud * a.P + p.k) * 3;
          p.v0 = __ldg(q); p.v1 = __ldg(q + 1); p.v2 = __ldg(q + 2);
        }
      }
      return p;
    };
    auto gather_begin = [&](int b, const Pre &p) {
      uint8_t *sA = smem + b * Cfg::A_BYTES;
      if (LEVEL == 1) {
        if (wg == 0) {
          uint4 row = make_uint4(0, 0, 0, 0);
          if (live) {
            row.x = pack_bf16(p.v0 - cx, p.v1 - cy);
            row.y = pack_bf16(p.v2 - cz, p.v3);
            row.z = pack_bf16(p.v4, p.v5);
          }
          *reinterpret_cast<uint4 *>(sA + Cfg::A_TAIL_OFF + tile_off(128, r, 0)) = row;  // columns 8..15 stay zero
        }
      } else {
        // columns [0,KF) = features of the neighbour, [KF,KF+3) = xyz - centre (weights permuted to match)
        constexpr int CH = Cfg::KF / 8;  // 16-byte chunks of the feature row
        const uint4 *f = reinterpret_cast<const uint4 *>(a.feat) + ((size_t)cloud * a.P + p.k) * CH;
#pragma unroll
        for (int q = 0; q < CH / NCG; ++q) {
          const int qq = wg * (CH / NCG) + q;
          cp_async16(sA + tile_off_sw128(128, r, qq * 8), f + qq, live ? 16u : 0u);
        }
        if (wg == NCG - 1) {
          uint4 row = make_uint4(0, 0, 0, 0);
          if (live) {
            row.x = pack_bf16(p.v0 - cx, p.v1 - cy);
            row.y = pack_bf16(p.v2 - cz, 0.f);
          }
          *reinterpret_cast<uint4 *>(sA + Cfg::A_TAIL_OFF + tile_off(128, r, 0)) = row;  // tail columns 8..15 stay zero
        }
      }
    };
    auto gather_end = [&]() {
      if (LEVEL != 1) cp_async_wait_all();
    };
    // relu(acc + shift) of layer L (1,2) of buffer b -> bf16 -> A buffer b (this thread: its row, half of the columns)
    auto epi12 = [&](int L, int b) {
      uint8_t *sA = smem + b * Cfg::A_BYTES;
      const uint32_t tb = tmem + b * Cfg::TMEM_COLS + lane_off;
      if (L == 1) {
        constexpr int NC = Cfg::N1 / NCG;  // 16 or 32 columns per thread
        const int c0 = wg * NC;
        epilogue_to_smem<NC>(tb + c0, sh1, sA, r, c0);
      } else {
        constexpr int NC = Cfg::N2 / NCG;
        const int c0 = wg * NC;
        epilogue_to_smem<NC>(tb + Cfg::N1 + c0, sh2, sA, r, c0);
      }
    };
    auto epi3 = [&](int b) {
      const uint32_t tb = tmem + b * Cfg::TMEM_COLS + lane_off + wg * (Cfg::N3 / NCG);
#pragma unroll
      for (int cc = 0; cc < Cfg::N3 / (32 * NCG); ++cc) {
        float v[32];
        tmem_ld32(tb + cc * 32, v);
        if constexpr (PACKED) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            runmax[cc * 16 + i] = __hmax2(runmax[cc * 16 + i], __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) runmaxf[cc * 32 + i] = fmaxf(runmaxf[cc * 32 + i], v[i]);
        }
      }
    };

    // ---- prologue: tiles s = 0 and s = 1 -------------------------------------------------------------------
    {
      const Pre p0 = preload(0), p1 = preload(1);
      gather_begin(0, p0);
      if (1 < NS) gather_begin(1, p1);
      gather_end();
      publish(1, 0);
      if (1 < NS) publish(1, 1);
    }
    for (int s = 0; s < NS; s += 2) {
      const bool has1 = s + 1 < NS;
      const Pre n0 = preload(s + 2), n1 = preload(s + 3);  // raw data of the next pair, consumed at the end of this one
#pragma unroll
      for (int L = 1; L <= 2; ++L) {
        wait_mma(0);
        epi12(L, 0);
        publish(L + 1, 0);
        if (has1) {
          wait_mma(1);
          epi12(L, 1);
          publish(L + 1, 1);
        }
      }
      // layer 3 done: A buffer free -> start the next gather, take the running max, hand the next tile to the tensor pipe
      wait_mma(0);
      if (s + 2 < NS) gather_begin(0, n0);
      epi3(0);
      if (s + 2 < NS) {
        gather_end();
        publish(1, 0);
      }
      if (has1) {
        wait_mma(1);
        if (s + 3 < NS) gather_begin(1, n1);
        epi3(1);
        if (s + 3 < NS) {
          gather_end();
          publish(1, 1);
        }
      }
    }
    // ---- finalize: relu(max + shift3) -> bf16 row of this centre -------------------------------------------
    if (live) {
      uint4 *o = reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(a.out) + (size_t)cg * Cfg::N3 +
                                            wg * (Cfg::N3 / NCG));
#pragma unroll
      for (int q = 0; q < Cfg::N3 / (8 * NCG); ++q) {
        uint32_t w[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int c = q * 8 + h * 2;
          const int gc = wg * (Cfg::N3 / NCG) + c;
          float2 m;
          if constexpr (PACKED) m = __bfloat1622float2(runmax[c >> 1]);
          else m = make_float2(runmaxf[c], runmaxf[c + 1]);
          w[h] = pack_bf16(fmaxf(m.x + sh3[gc], 0.f), fmaxf(m.y + sh3[gc + 1], 0.f));
        }
        o[q] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    // the next super-tile's first MMA overwrites TMEM that this one's last epilogue read: ordered by the a_ready arrival
    // of the next prologue (publish() carries the tcgen05 fence) / by the CTA barrier of the next prologue
  }
  fence_before_sync();
  __syncthreads();
  if (tmem_owner) tmem_dealloc<2 * Cfg::TMEM_COLS>(tmem);
}

using Sa1 = SaCfg<3, 16, 64, 64, 128, 32, 2>;   // 256 + 32 threads, 2 CTAs / SM
using Sa2 = SaCfg<128, 144, 128, 128, 256, 16, 4>;  // 512 + 32 threads, 1 CTA / SM

template <class Cfg, int LEVEL>
int launch_sa(const SaMlpArgs &a, cudaStream_t st) {
  auto kern = sa_mlp_kernel<Cfg, LEVEL>;
  int rc = sv::cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  if (rc) return rc;
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  rc = sv::cuda_status(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Cfg::NTHREADS_ALL, Cfg::SMEM_BYTES));
  if (rc) return rc;
  if (per_sm < 1) per_sm = 1;
  if (per_sm * 2 * Cfg::TMEM_COLS > 512) per_sm = 512 / (2 * Cfg::TMEM_COLS);  // TMEM columns are not part of the occupancy query
  const int n_super = (a.B * Cfg::CPC + 127) / 128;
  int grid = sms * per_sm;
  if (grid > n_super) grid = n_super;
  kern<<<grid, Cfg::NTHREADS_ALL, Cfg::SMEM_BYTES, st>>>(a);
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
