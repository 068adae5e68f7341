// Fused set-abstraction sampling: furthest point sampling + ball query (+ the same for the next
// set-abstraction level on the sampled centres) in ONE pass over the cloud, one warp per cloud.
//
// Replaces, for the SA modules of the GPS object encoder, the reference call sequence
//   furthest_point_sample -> gather_operation -> ball_query            (pointnet2_modules.py:54-58,
//   pointnet2_utils.py:331) x 2 levels
// and produces bit-identical indices (tests/test_pointops_gpu.py).
//
// Key observation: FPS iteration j evaluates d(k, centre_j) for every point k — exactly the
// distances the ball query of centre j needs (same operands, same fma order: (a-b)^2 terms are
// sign-symmetric, so sqdist(point,centre) == sqdist(centre,point) bit for bit).  The ball query is
// therefore a by-product: one compare + one warp ballot per register slot, the ballot words are
// the hit bitmaps in point-index order (the slot layout makes slot i == 32 consecutive indices in
// bit-reversed lane order), and the "first nsample in index order" rule of
// ball_query_gpu.cu:27-42 becomes a bit-scan of <= 32 words per centre.
// Distances use packed fp32x2 instructions (FADD2/FMUL2/FFMA2, IEEE-rn per element, sm_100+).
#include "svcommon.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;
typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
  u64 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}
__device__ __forceinline__ bool mag_skipped(float mag) { return mag < __uint_as_float(0x3A83126Fu); }
__device__ __forceinline__ int brevn(int v, int n) { return n ? (int)(__brev((unsigned)v) >> (32 - n)) : 0; }

__device__ __forceinline__ void stage_floats(float *dst, const float *__restrict__ src, int count, int tid,
                                             int nthreads) {
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int n4 = count >> 2;
    const float4 *s4 = reinterpret_cast<const float4 *>(src);
    float4 *d4 = reinterpret_cast<float4 *>(dst);
    for (int i = tid; i < n4; i += nthreads) d4[i] = __ldg(s4 + i);
    for (int i = (n4 << 2) + tid; i < count; i += nthreads) dst[i] = __ldg(src + i);
  } else {
    for (int i = tid; i < count; i += nthreads) dst[i] = __ldg(src + i);
  }
}

// log-depth maximum of v[0..N) with 3-input max (FMNMX3 on sm_100a)
template <int N>
__device__ __forceinline__ float max_tree(const float *v) {
  if constexpr (N == 1) {
    return v[0];
  } else if constexpr (N == 2) {
    return fmaxf(v[0], v[1]);
  } else if constexpr (N == 3) {
    return fmaxf(fmaxf(v[0], v[1]), v[2]);
  } else {
    constexpr int A = (N + 2) / 3, B2 = (N - A + 1) / 2, C = N - A - B2;
    return fmaxf(fmaxf(max_tree<A>(v), max_tree<B2>(v + A)), max_tree<C>(v + A + B2));
  }
}

struct SaSampleParams {
  const float *xyz;  // (B,N,3)
  int N, m;          // 32 <= N <= 1024
  int BS, lgBS, Qmax, spt;
  float radius;
  int nsample;
  int *fps_idx;    // (B,m)
  float *new_xyz;  // (B,m,3)
  int *ball_idx;   // (B,m,nsample)
  // optional second level on the m == 32 sampled centres (m2 == 0: off)
  int m2;
  float radius_2;
  int nsample2;
  int *fps_idx2;    // (B,m2)
  float *new_xyz2;  // (B,m2,3)
  int *ball_idx2;   // (B,m2,nsample2)
};

constexpr float FAR_AWAY = 1e18f;  // coordinates of non-existent slots: never within any radius

template <int SPT>
__global__ void __launch_bounds__(32) sa_sample_kernel(const SaSampleParams p) {
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x;
  const int b = blockIdx.x;
  const int N = p.N, m = p.m, nsample = p.nsample;
  const int n3 = 3 * N;
  const int rs = nsample | 1;
  float *pts = sm;                                             // [3N] AoS copy of the cloud
  int *kbase = reinterpret_cast<int *>(sm + ((n3 + 3) & ~3));  // [32]
  int *slot_of_chunk = kbase + 32;                             // [32] slot holding points [32c, 32c+32)
  unsigned *W = reinterpret_cast<unsigned *>(slot_of_chunk + 32);  // [32][33] hit words: row = centre, col = slot
  float *rec = reinterpret_cast<float *>(W + 32 * 33);         // [32] winner lane's running distances
  // [32][rs] compaction staging; with a single batch of centres (m <= 32) the cloud copy is dead by then (the last FPS
  // iteration selects nothing), so the staging rows alias it: 17 KB instead of 21 KB per cloud = 12 instead of 10 clouds per SM
  const bool alias_stage = m <= 32 && (size_t)32 * rs <= (size_t)n3;
  int *stage = alias_stage ? reinterpret_cast<int *>(pts) : reinterpret_cast<int *>(rec + 32);

  stage_floats(pts, p.xyz + (size_t)b * n3, n3, lane, 32);
  if (lane < p.spt) {
    const int u = lane / p.Qmax, r = lane - u * p.Qmax;
    const int kb = r * p.BS + (brevn(u, p.lgBS - 5) << 5);
    kbase[lane] = kb;
    slot_of_chunk[kb >> 5] = lane;
  }
  __syncwarp();

  // lane L holds, in slot i, point kbase[i] + L; its position in the reference's tie-break order is
  // brev5(L) (bit-reversed thread id), so ballots come out in point-index order and only the argmax
  // candidate key carries the bit reversal.
  const int lanebits = lane;
  constexpr int NP = (SPT + 1) / 2;  // register pairs
  u64 px[NP], py[NP], pz[NP];
  float pt[2 * NP];
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    float cx[2], cy[2], cz[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = 2 * q + h;
      float x = FAR_AWAY, y = FAR_AWAY, z = FAR_AWAY, t = -2.0f;  // inert for FPS, never a ball-query hit
      if (i < p.spt) {
        const int k = kbase[i] + lanebits;
        if (k < N) {
          x = pts[3 * k + 0];
          y = pts[3 * k + 1];
          z = pts[3 * k + 2];
          const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
          if (!mag_skipped(mag)) t = 1e10f;  // skipped points stay inert for FPS but are ball-query candidates
        }
      }
      cx[h] = x; cy[h] = y; cz[h] = z;
      pt[i] = t;
    }
    px[q] = pack2(cx[0], cx[1]);
    py[q] = pack2(cy[0], cy[1]);
    pz[q] = pack2(cz[0], cz[1]);
  }

  const float r2 = __fmul_rn(p.radius, p.radius);
  const int nchunks = (N + 31) >> 5;
  int old = 0;
  float x1 = pts[0], y1 = pts[1], z1 = pts[2];
  int my_idx = 0;
  float mcx = 0.f, mcy = 0.f, mcz = 0.f;
  for (int j = 0; j < m; ++j) {
    const int jr = j & 31;
    const u64 X1 = pack2(x1, x1), Y1 = pack2(y1, y1), Z1 = pack2(z1, z1);
    unsigned *wrow = W + jr * 33;
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      const u64 dx = sub2(px[q], X1), dy = sub2(py[q], Y1), dz = sub2(pz[q], Z1);
      const u64 dd = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
      float d[2];
      unpack2(dd, d[0], d[1]);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = 2 * q + h;
        if (i < SPT) {
          // bit L of the ballot <-> point kbase[i] + L; lane 0 files the (warp-uniform) word under its slot
          const unsigned w = __ballot_sync(FULL, d[h] < r2);
          if (lane == 0) wrow[i] = w;
          pt[i] = fminf(d[h], pt[i]);
        }
      }
    }
    if (lane == jr) {
      my_idx = old;
      mcx = x1; mcy = y1; mcz = z1;
    }
    if (jr == 31 || j == m - 1) {
      const int base = j & ~31;
      const int nrows = j - base + 1;
      const size_t o = (size_t)b * m + base;
      if (lane < nrows) {
        p.fps_idx[o + lane] = my_idx;
        float *q = p.new_xyz + (o + lane) * 3;
        q[0] = mcx; q[1] = mcy; q[2] = mcz;
      }
      __syncwarp();
      // lane == centre: walk its hit string (one 32-bit word per 32-point chunk) in index order and
      // emit exactly `nsample` entries: the hits, then the first hit as padding (0 when there is none).
      {
        const unsigned *wr = W + lane * 33;
        unsigned nz = 0u;  // chunks with at least one hit
        for (int c = 0; c < nchunks; ++c) nz |= (wr[slot_of_chunk[c]] != 0u ? 1u : 0u) << c;
        if (lane >= nrows) nz = 0u;
        int *row = stage + lane * rs;
        unsigned w = 0u;
        int cb = 0, first = 0;
        for (int s = 0; s < nsample; ++s) {
          if (w == 0u && nz != 0u) {
            const int c = __ffs(nz) - 1;
            nz &= nz - 1u;
            w = wr[slot_of_chunk[c]];
            cb = c << 5;
          }
          int v = first;
          if (w != 0u) {
            v = cb + __ffs(w) - 1;
            w &= w - 1u;
          }
          if (s == 0) first = v;
          row[s] = v;
        }
      }
      __syncwarp();
      {
        const int total = nrows * nsample;
        int *out = p.ball_idx + o * (size_t)nsample;
        if (nsample == 32) {
          for (int e = lane; e < total; e += 32) out[e] = stage[(e >> 5) * rs + lane];
        } else {
          for (int e = lane; e < total; e += 32) {
            const int r = e / nsample, s2 = e - r * nsample;
            out[e] = stage[r * rs + s2];
          }
        }
      }
      __syncwarp();
    }
    if (j < m - 1) {
      // lane maximum of the running distances (log-depth tree, 3-input max), then the warp maximum;
      // the winner among equal distances is the lowest (bit-reversed lane, slot): CREDUX.MIN picks the
      // lane, which publishes its 32 slot values so that a ballot finds its first slot equal to the max.
      const float best = max_tree<SPT>(pt);
      const unsigned key = best < 0.f ? 0u : __float_as_uint(best) + 1u;
      const unsigned M = __reduce_max_sync(FULL, key);
      if (M == 0u) {
        old = 0;  // every reference thread reported (best=-1, besti=0)
      } else {
        const unsigned rk = (unsigned)brevn(lane, 5);
        const unsigned wl = __reduce_min_sync(FULL, key == M ? rk : 0xffffffffu);  // winning lane's rank
        if (rk == wl) {
#pragma unroll
          for (int i = 0; i < SPT; ++i) rec[i] = pt[i];
        }
        __syncwarp();
        const float mine = lane < SPT ? rec[lane] : -3.0f;
        const unsigned eq = __ballot_sync(FULL, __float_as_uint(mine) + 1u == M);
        __syncwarp();
        old = kbase[__ffs(eq) - 1] + brevn((int)wl, 5);
      }
      x1 = pts[3 * old + 0];
      y1 = pts[3 * old + 1];
      z1 = pts[3 * old + 2];
    }
  }

  // ---------------- second level: FPS + ball query over the 32 centres just sampled ----------------
  if (p.m2 > 0) {
    // level-2 cloud: point k = centre k, held by lane k.  Reference geometry for n = 32: BS = 32, one
    // point per thread, ties go to the smallest bit-reversed thread id.
    const float qx = mcx, qy = mcy, qz = mcz;
    const float mag = __fmaf_rn(qz, qz, __fmaf_rn(qx, qx, __fmul_rn(qy, qy)));
    float t = mag_skipped(mag) ? -2.0f : 1e10f;
    const float r22 = __fmul_rn(p.radius_2, p.radius_2);
    const int ns2 = p.nsample2;
    int old2 = 0;
    for (int j = 0; j < p.m2; ++j) {
      const float c1 = __shfl_sync(FULL, qx, old2), c2 = __shfl_sync(FULL, qy, old2), c3 = __shfl_sync(FULL, qz, old2);
      const float d = sqdist(qx, qy, qz, c1, c2, c3);
      const unsigned word = __ballot_sync(FULL, d < r22);  // bit k <-> point k
      const size_t o = (size_t)b * p.m2 + j;
      if (lane == 0) {
        p.fps_idx2[o] = old2;
        float *q = p.new_xyz2 + o * 3;
        q[0] = c1; q[1] = c2; q[2] = c3;
      }
      const int cnt = __popc(word);
      const int first = cnt ? __ffs(word) - 1 : 0;
      for (int s = lane; s < ns2; s += 32)
        p.ball_idx2[o * ns2 + s] = s < cnt ? (int)__fns(word, 0, s + 1) : first;
      if (j < p.m2 - 1) {
        const float d2 = fminf(d, t);
        t = d2;
        const unsigned key = d2 < 0.f ? 0u : __float_as_uint(d2) + 1u;
        const unsigned M = __reduce_max_sync(FULL, key);
        if (M == 0u) {
          old2 = 0;
        } else {
          const unsigned w = __reduce_min_sync(FULL, key == M ? (unsigned)brevn(lane, 5) : 0xffffffffu);
          old2 = brevn((int)w, 5);
        }
      }
    }
  }
}

template <int SPT>
int launch(const SaSampleParams &p, int B, cudaStream_t st) {
  const bool alias_stage = p.m <= 32 && (size_t)32 * (p.nsample | 1) <= (size_t)3 * p.N;
  const size_t smem = (size_t)((3 * p.N + 3) & ~3) * 4 + 64 * 4 + 32 * 33 * 4 + 32 * 4 +
                      (alias_stage ? 0 : (size_t)32 * (p.nsample | 1) * 4);
  auto kern = sa_sample_kernel<SPT>;
  if (smem > 48 * 1024) {
    int rc = sv::cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (rc) return rc;
  }
  kern<<<B, 32, smem, st>>>(p);
  return sv::after_launch();
}

}  // namespace

extern "C" int sv_sa_sample_f32(const float *xyz, int B, int N, int m, float radius, int nsample, int *fps_idx,
                                float *new_xyz, int *ball_idx, int m2, float radius_2, int nsample2, int *fps_idx2,
                                float *new_xyz2, int *ball_idx2, void *stream) {
  if (B < 0 || N < 32 || N > 1024 || m < 1 || nsample < 1 || nsample > 256) return SV_ERR_INVALID_ARG;
  if (m2 < 0 || (m2 > 0 && (m != 32 || nsample2 < 1 || nsample2 > 256))) return SV_ERR_INVALID_ARG;
  if (B == 0) return SV_OK;
  if (!xyz || !fps_idx || !new_xyz || !ball_idx) return SV_ERR_INVALID_ARG;
  if (m2 > 0 && (!fps_idx2 || !new_xyz2 || !ball_idx2)) return SV_ERR_INVALID_ARG;
  SaSampleParams p;
  p.xyz = xyz; p.N = N; p.m = m;
  p.BS = sv::ref_opt_n_threads(N);
  p.lgBS = 0;
  while ((1 << p.lgBS) < p.BS) ++p.lgBS;
  p.Qmax = (N + p.BS - 1) / p.BS;
  p.spt = (p.BS >> 5) * p.Qmax;
  p.radius = radius; p.nsample = nsample;
  p.fps_idx = fps_idx; p.new_xyz = new_xyz; p.ball_idx = ball_idx;
  p.m2 = m2; p.radius_2 = radius_2; p.nsample2 = nsample2;
  p.fps_idx2 = fps_idx2; p.new_xyz2 = new_xyz2; p.ball_idx2 = ball_idx2;
  cudaStream_t st = (cudaStream_t)stream;
  if (p.spt <= 1) return launch<1>(p, B, st);
  if (p.spt <= 2) return launch<2>(p, B, st);
  if (p.spt <= 4) return launch<4>(p, B, st);
  if (p.spt <= 8) return launch<8>(p, B, st);
  if (p.spt <= 16) return launch<16>(p, B, st);
  if (p.spt <= 32) return launch<32>(p, B, st);
  return SV_ERR_INVALID_ARG;
}
