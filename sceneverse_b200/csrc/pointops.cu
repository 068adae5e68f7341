// libsvpointops: sm_100a kernels for the PointNet++ point operators behind the C-ABI in
// include/svpointops.h (the drop-in for the reference's `pointnet2._ext`,
// reference: modules/third_party/pointnet2/_ext_src/src/*.cu).
//
// Bit-exactness rules shared by every distance below (PTX-verified contraction order of the
// reference sources): d2 = fma(dz,dz, fma(dx,dx, dy*dy)).
//
// Furthest point sampling keeps the whole per-cloud state on chip: the cloud is staged once
// through shared memory with 128-bit coalesced loads, each lane then owns SPT "slots" (points) in
// REGISTERS (x,y,z and the running min-distance), so one FPS iteration is a register-only sweep
// followed by two warp REDUX instructions.  The reference's winner among equal distances is
// (max dist, then min bit-reversed (k mod BS), then min k) with BS = the reference block size
// (sampling_gpu.cu:59-65,115-168); slots are laid out so that this order is simply "lowest
// (lane, slot)" and no tree has to be replayed.
#include <math.h>
#include <stdio.h>

#include "svcommon.h"

namespace sv {
std::atomic<unsigned long long> g_launches{0};
thread_local int t_last_cuda_error = 0;

int ref_opt_n_threads(int work_size) {
  const int pow_2 = (int)(std::log(static_cast<double>(work_size)) / std::log(2.0));
  int v = 1 << pow_2;
  if (v > 512) v = 512;
  if (v < 1) v = 1;
  return v;
}
}  // namespace sv

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// sampling_gpu.cu:100-101 skips a point iff (double)mag <= 1e-3.  The largest fp32 value whose
// double is <= 1e-3 is 0x3A83126E, so the test is mag < 0x3A83126F in fp32 (NaN -> not skipped in
// both forms); tests/test_oracle.py checks the two boundary values.
__device__ __forceinline__ bool mag_skipped(float mag) { return mag < __uint_as_float(0x3A83126Fu); }

// reverse the low n bits of v (n may be 0)
__device__ __forceinline__ int brevn(int v, int n) { return n ? (int)(__brev((unsigned)v) >> (32 - n)) : 0; }

// Coalesced copy of `count` floats global -> shared by `nthreads` threads (128-bit when aligned).
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

// ------------------------------------------------------------------------------------------------
// Ball query core, one warp, lane == centre (exactly the reference's thread mapping,
// ball_query_gpu.cu:24-43, but with the points broadcast from shared memory and the result rows
// staged in shared memory so that the global store is one coalesced stream).
// pts: AoS xyz of points [k0, k0+count) of the cloud; row: this lane's staging row.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bq_scan_tile(const float *__restrict__ pts, int k0, int count, float cx, float cy,
                                             float cz, float r2, int nsample, int *__restrict__ row, int &cnt) {
  for (int kb = 0; kb < count; kb += 64) {
    if (__all_sync(FULL, cnt >= nsample)) break;
    const int ke = min(kb + 64, count);
#pragma unroll 4
    for (int k = kb; k < ke; ++k) {
      const float x = pts[3 * k + 0], y = pts[3 * k + 1], z = pts[3 * k + 2];
      const float d2 = sqdist(cx, cy, cz, x, y, z);
      if (d2 < r2 && cnt < nsample) {
        row[cnt] = k0 + k;
        ++cnt;
      }
    }
  }
}

// rows [0,nrows) of this warp's staging area -> idx rows (contiguous in global), padding each row
// with its first hit (ball_query_gpu.cu:34-38) or zeros when there was none (ball_query.cpp:19-21).
__device__ __forceinline__ void bq_flush_rows(const int *__restrict__ stage, const int *__restrict__ cnts, int rs,
                                              int nrows, int nsample, int *__restrict__ out, int lane) {
  const int total = nrows * nsample;
  for (int e = lane; e < total; e += 32) {
    const int r = e / nsample, s = e - r * nsample;
    const int c = cnts[r];
    out[e] = s < c ? stage[r * rs + s] : (c > 0 ? stage[r * rs] : 0);
  }
}

// ------------------------------------------------------------------------------------------------
// FPS, one warp per cloud, N <= 1024 (32 lanes x SPT register slots).  Optionally fused with the
// ball query of the sampled centres (FUSE_BQ).
// Geometry (host-computed): BS = reference block size, lgBS, Qmax = ceil(N/BS),
// lgT = log2(min(32,BS)) active lanes, spt = (BS >> lgT) * Qmax <= SPT.
// Slot (lane L, i) holds point k = (i % Qmax) * BS + (brev(i / Qmax) << lgT) + brev_lgT(L).
// ------------------------------------------------------------------------------------------------
template <int SPT, bool FUSE_BQ>
__global__ void __launch_bounds__(32)
fps_warp_kernel(const float *__restrict__ xyz, int N, int m, int BS, int lgBS, int Qmax, int lgT, int spt,
                int *__restrict__ idx, float *__restrict__ new_xyz, float radius, int nsample,
                int *__restrict__ ball_idx) {
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x;
  const int b = blockIdx.x;
  const int n3 = 3 * N;
  float *pts = sm;                                               // [3N] AoS copy of the cloud
  int *kbase = reinterpret_cast<int *>(sm + ((n3 + 3) & ~3));    // [32]
  int *bq_cnt = kbase + 32;                                      // [32]            (FUSE_BQ)
  int *bq_stage = bq_cnt + 32;                                   // [32][nsample|1] (FUSE_BQ)

  stage_floats(pts, xyz + (size_t)b * n3, n3, lane, 32);
  if (lane < spt) {
    const int u = lane / Qmax, r = lane - u * Qmax;
    kbase[lane] = r * BS + (brevn(u, lgBS - lgT) << lgT);
  }
  __syncwarp();

  const bool lane_on = lane < (1 << lgT);
  const int lanebits = brevn(lane, lgT);
  float px[SPT], py[SPT], pz[SPT], pt[SPT];
#pragma unroll
  for (int i = 0; i < SPT; ++i) {
    float x = 0.f, y = 0.f, z = 0.f, t = -2.0f;  // inert slot: min(d,-2) never beats best >= -1
    if (lane_on && i < spt) {
      const int k = kbase[i] + lanebits;
      if (k < N) {
        x = pts[3 * k + 0];
        y = pts[3 * k + 1];
        z = pts[3 * k + 2];
        const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
        if (!mag_skipped(mag)) t = 1e10f;  // sampling_gpu.cu:100-101
      }
    }
    px[i] = x; py[i] = y; pz[i] = z; pt[i] = t;
  }

  int old = 0;
  float x1 = pts[0], y1 = pts[1], z1 = pts[2];
  int my_idx = 0;
  float cx = 0.f, cy = 0.f, cz = 0.f;
  const float r2 = __fmul_rn(radius, radius);
  const int rs = nsample | 1;
  for (int j = 0; j < m; ++j) {
    if (j > 0) {
      float best = -1.0f;
      int besti = 0;
#pragma unroll
      for (int i = 0; i < SPT; ++i) {
        const float d = sqdist(px[i], py[i], pz[i], x1, y1, z1);
        const float d2 = fminf(d, pt[i]);
        pt[i] = d2;
        const bool p = d2 > best;
        besti = p ? i : besti;
        best = p ? d2 : best;
      }
      const unsigned key = best < 0.f ? 0u : __float_as_uint(best) + 1u;
      const unsigned M = __reduce_max_sync(FULL, key);
      if (M == 0u) {
        old = 0;  // every reference thread reported (best=-1, besti=0)
      } else {
        const unsigned cand = key == M ? (unsigned)((lane << 8) | besti) : 0xffffffffu;
        const unsigned w = __reduce_min_sync(FULL, cand);
        old = kbase[w & 255u] + brevn((int)(w >> 8), lgT);
      }
      x1 = pts[3 * old + 0];
      y1 = pts[3 * old + 1];
      z1 = pts[3 * old + 2];
    }
    if (lane == (j & 31)) {
      my_idx = old;
      cx = x1; cy = y1; cz = z1;
    }
    if ((j & 31) == 31 || j == m - 1) {
      const int base = j & ~31;
      const int nrows = j - base + 1;
      const size_t o = (size_t)b * m + base;
      if (lane < nrows) {
        idx[o + lane] = my_idx;
        if (new_xyz != nullptr) {
          float *q = new_xyz + (o + lane) * 3;
          q[0] = cx; q[1] = cy; q[2] = cz;
        }
      }
      if (FUSE_BQ) {
        int cnt = lane < nrows ? 0 : nsample;  // idle lanes count as full so the scan can exit early
        bq_scan_tile(pts, 0, N, cx, cy, cz, lane < nrows ? r2 : -1.0f, nsample, bq_stage + lane * rs, cnt);
        bq_cnt[lane] = cnt;
        __syncwarp();
        bq_flush_rows(bq_stage, bq_cnt, rs, nrows, nsample, ball_idx + o * (size_t)nsample, lane);
        __syncwarp();
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// FPS, one CTA of T=256 threads per cloud, 1024 < N <= 8192 (BS = 512): thread t owns SPT slots in
// registers, spt = 2*Qmax; slot (t,i) holds k = (i % Qmax)*512 + (brev_1(i / Qmax) << 8) + brev_8(t).
// ------------------------------------------------------------------------------------------------
template <int SPT>
__global__ void __launch_bounds__(256)
fps_block_kernel(const float *__restrict__ xyz, int N, int m, int Qmax, int spt, int *__restrict__ idx,
                 float *__restrict__ new_xyz) {
  constexpr int T = 256, LGT = 8, LGBS = 9, BS = 512, W = T / 32;
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const int n3 = 3 * N;
  float *pts = sm;
  int *kbase = reinterpret_cast<int *>(sm + ((n3 + 3) & ~3));  // [SPT]
  unsigned *red = reinterpret_cast<unsigned *>(kbase + SPT);   // [2][W][2]

  stage_floats(pts, xyz + (size_t)b * n3, n3, tid, T);
  if (tid < spt) {
    const int u = tid / Qmax, r = tid - u * Qmax;
    kbase[tid] = r * BS + (brevn(u, LGBS - LGT) << LGT);
  }
  __syncthreads();

  const int tbits = brevn(tid, LGT);
  float px[SPT], py[SPT], pz[SPT], pt[SPT];
#pragma unroll
  for (int i = 0; i < SPT; ++i) {
    float x = 0.f, y = 0.f, z = 0.f, t = -2.0f;
    if (i < spt) {
      const int k = kbase[i] + tbits;
      if (k < N) {
        x = pts[3 * k + 0];
        y = pts[3 * k + 1];
        z = pts[3 * k + 2];
        const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
        if (!mag_skipped(mag)) t = 1e10f;
      }
    }
    px[i] = x; py[i] = y; pz[i] = z; pt[i] = t;
  }

  int old = 0;
  float x1 = pts[0], y1 = pts[1], z1 = pts[2];
  for (int j = 0; j < m; ++j) {
    if (j > 0) {
      float best = -1.0f;
      int besti = 0;
#pragma unroll
      for (int i = 0; i < SPT; ++i) {
        const float d = sqdist(px[i], py[i], pz[i], x1, y1, z1);
        const float d2 = fminf(d, pt[i]);
        pt[i] = d2;
        const bool p = d2 > best;
        besti = p ? i : besti;
        best = p ? d2 : best;
      }
      const unsigned key = best < 0.f ? 0u : __float_as_uint(best) + 1u;
      const unsigned Mw = __reduce_max_sync(FULL, key);
      const unsigned cand = (key == Mw && Mw != 0u) ? (unsigned)((tid << 8) | besti) : 0xffffffffu;
      const unsigned ww = __reduce_min_sync(FULL, cand);
      unsigned *rb = red + (j & 1) * (2 * W);
      if (lane == 0) {
        rb[2 * warp] = Mw;
        rb[2 * warp + 1] = ww;
      }
      __syncthreads();
      const unsigned k2 = lane < W ? rb[2 * lane] : 0u;
      const unsigned c2 = lane < W ? rb[2 * lane + 1] : 0xffffffffu;
      const unsigned M = __reduce_max_sync(FULL, k2);
      if (M == 0u) {
        old = 0;
      } else {
        const unsigned w = __reduce_min_sync(FULL, k2 == M ? c2 : 0xffffffffu);
        old = kbase[w & 255u] + brevn((int)(w >> 8), LGT);
      }
      x1 = pts[3 * old + 0];
      y1 = pts[3 * old + 1];
      z1 = pts[3 * old + 2];
    }
    if (tid == 0) {
      idx[(size_t)b * m + j] = old;
      if (new_xyz != nullptr) {
        float *q = new_xyz + ((size_t)b * m + j) * 3;
        q[0] = x1; q[1] = y1; q[2] = z1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// FPS, any N (BS = 512 for N >= 512): one CTA of 1024 threads per cloud, running min-distance in a
// global scratch row (zero-copy of the reference's `temp`, sampling.cpp:74-76), xyz streamed from
// L2.  The reference order is applied explicitly: key = (dist, then min rank(k)) with
// rank(k) = brev_lg(k mod BS) * Qmax + k / BS.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
fps_generic_kernel(const float *__restrict__ xyz, int N, int m, int BS, int lgBS, int Qmax,
                   float *__restrict__ temp, int *__restrict__ idx, float *__restrict__ new_xyz) {
  constexpr int T = 1024, W = T / 32;
  __shared__ unsigned long long red[2][W];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const float *pts = xyz + (size_t)b * 3 * N;
  float *tmp = temp + (size_t)b * N;
  // init: -2 marks points the reference never touches (|p|^2 <= 1e-3), 1e10 otherwise
  for (int k = tid; k < N; k += T) {
    const float x = pts[3 * k], y = pts[3 * k + 1], z = pts[3 * k + 2];
    const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
    tmp[k] = mag_skipped(mag) ? -2.0f : 1e10f;
  }
  __syncthreads();
  int old = 0;
  for (int j = 0; j < m; ++j) {
    const float x1 = pts[3 * old], y1 = pts[3 * old + 1], z1 = pts[3 * old + 2];
    if (tid == 0) {
      idx[(size_t)b * m + j] = old;
      if (new_xyz != nullptr) {
        float *q = new_xyz + ((size_t)b * m + j) * 3;
        q[0] = x1; q[1] = y1; q[2] = z1;
      }
    }
    if (j == m - 1) break;
    // key: high 32 = float bits + 1 (0 = nothing valid), low 32 = ~rank  -> max wins
    unsigned long long best = 0ull;
    for (int k = tid; k < N; k += T) {
      const float t = tmp[k];
      if (t < 0.f) continue;
      const float d = sqdist(pts[3 * k], pts[3 * k + 1], pts[3 * k + 2], x1, y1, z1);
      const float d2 = fminf(d, t);
      tmp[k] = d2;
      const unsigned rank = (unsigned)brevn(k & (BS - 1), lgBS) * (unsigned)Qmax + (unsigned)(k >> lgBS);
      const unsigned long long key =
          ((unsigned long long)(__float_as_uint(d2) + 1u) << 32) | (unsigned long long)(0xffffffffu - rank);
      best = key > best ? key : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(FULL, best, o);
      best = other > best ? other : best;
    }
    if (lane == 0) red[j & 1][warp] = best;
    __syncthreads();
    unsigned long long v = red[j & 1][lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(FULL, v, o);
      v = other > v ? other : v;
    }
    if (v == 0ull) {
      old = 0;
    } else {
      const unsigned rank = 0xffffffffu - (unsigned)(v & 0xffffffffull);
      const unsigned br = rank / (unsigned)Qmax, q = rank - br * (unsigned)Qmax;
      old = (int)(q << lgBS) + brevn((int)br, lgBS);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Ball query, standalone: CTA = W warps, each warp owns 32 centres (lane == centre); the cloud
// streams through shared memory in tiles of TILE points shared by all warps of the CTA.
// ------------------------------------------------------------------------------------------------
constexpr int BQ_TILE = 2048;

__global__ void __launch_bounds__(256)
ball_query_kernel(const float *__restrict__ new_xyz, const float *__restrict__ xyz, int N, int M, float radius,
                  int nsample, int *__restrict__ idx) {
  extern __shared__ __align__(16) float sm[];
  const int W = blockDim.x >> 5;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunks = (M + 32 * W - 1) / (32 * W);
  const int b = blockIdx.x / chunks, chunk = blockIdx.x - b * chunks;
  const int rs = nsample | 1;
  float *pts = sm;                                                  // [3*BQ_TILE]
  int *cnts = reinterpret_cast<int *>(sm + 3 * BQ_TILE) + warp * 32;  // [W][32]
  int *stage = reinterpret_cast<int *>(sm + 3 * BQ_TILE) + W * 32 + warp * 32 * rs;  // [W][32][rs]

  const int c0 = (chunk * W + warp) * 32;  // first centre of this warp
  const int j = c0 + lane;
  const bool has = j < M;
  float cx = 0.f, cy = 0.f, cz = 0.f;
  if (has) {
    const float *q = new_xyz + ((size_t)b * M + j) * 3;
    cx = q[0]; cy = q[1]; cz = q[2];
  }
  const float r2 = has ? __fmul_rn(radius, radius) : -1.0f;
  int cnt = has ? 0 : nsample;  // idle lanes count as full so the scan can exit early
  const float *src = xyz + (size_t)b * 3 * N;
  for (int k0 = 0; k0 < N; k0 += BQ_TILE) {
    const int count = min(BQ_TILE, N - k0);
    __syncthreads();
    stage_floats(pts, src + (size_t)3 * k0, 3 * count, tid, blockDim.x);
    __syncthreads();
    if (c0 < M) bq_scan_tile(pts, k0, count, cx, cy, cz, r2, nsample, stage + lane * rs, cnt);
  }
  if (c0 < M) {
    cnts[lane] = cnt;
    __syncwarp();
    bq_flush_rows(stage, cnts, rs, min(32, M - c0), nsample, idx + ((size_t)b * M + c0) * nsample, lane);
  }
}

// ------------------------------------------------------------------------------------------------
// group_points / gather_points: out[b,c,e] = points[b,c,idx[b,e]], e over NP*NS (contiguous in idx
// and out).  Each thread keeps its (up to 4) indices in registers and loops over the channels, so
// idx is read once and every store is a coalesced 128-bit stream.
// ------------------------------------------------------------------------------------------------
template <bool VEC4>
__global__ void __launch_bounds__(256)
group_points_kernel(const float *__restrict__ points, const int *__restrict__ idx, int C, int N, int E,
                    float *__restrict__ out) {
  const int b = blockIdx.y;
  const float *p = points + (size_t)b * C * N;
  const int *ix = idx + (size_t)b * E;
  float *o = out + (size_t)b * C * E;
  if (VEC4) {
    const int e4 = blockIdx.x * blockDim.x + threadIdx.x;
    if (e4 * 4 >= E) return;
    const int4 ii = __ldg(reinterpret_cast<const int4 *>(ix) + e4);
    for (int c = 0; c < C; ++c) {
      const float *row = p + (size_t)c * N;
      float4 v;
      v.x = __ldg(row + ii.x);
      v.y = __ldg(row + ii.y);
      v.z = __ldg(row + ii.z);
      v.w = __ldg(row + ii.w);
      reinterpret_cast<float4 *>(o + (size_t)c * E)[e4] = v;
    }
  } else {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int ii = __ldg(ix + e);
    for (int c = 0; c < C; ++c) o[(size_t)c * E + e] = __ldg(p + (size_t)c * N + ii);
  }
}

// grad: grad_points[b,c,idx[b,e]] += grad_out[b,c,e]  (fp32 RED, same as the reference's atomicAdd)
__global__ void __launch_bounds__(256)
group_points_grad_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx, int C, int N, int E,
                         float *__restrict__ grad_points) {
  const int b = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int ii = __ldg(idx + (size_t)b * E + e);
  const float *g = grad_out + (size_t)b * C * E;
  float *gp = grad_points + (size_t)b * C * N;
  for (int c = 0; c < C; ++c) atomicAdd(gp + (size_t)c * N + ii, __ldg(g + (size_t)c * E + e));
}

// ------------------------------------------------------------------------------------------------
// three_nn: thread per unknown point, known points broadcast from shared-memory tiles; best
// distances kept in double with strict '<' exactly as interpolate_gpu.cu:27-49.
// ------------------------------------------------------------------------------------------------
constexpr int NN_TILE = 2048;
__global__ void __launch_bounds__(256)
three_nn_kernel(const float *__restrict__ unknown, const float *__restrict__ known, int n, int m,
                float *__restrict__ dist2, int *__restrict__ idx) {
  __shared__ __align__(16) float pts[3 * NN_TILE];
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const bool has = j < n;
  float ux = 0.f, uy = 0.f, uz = 0.f;
  if (has) {
    const float *u = unknown + ((size_t)b * n + j) * 3;
    ux = u[0]; uy = u[1]; uz = u[2];
  }
  double best1 = 1e40, best2 = 1e40, best3 = 1e40;
  int besti1 = 0, besti2 = 0, besti3 = 0;
  const float *src = known + (size_t)b * 3 * m;
  for (int k0 = 0; k0 < m; k0 += NN_TILE) {
    const int count = min(NN_TILE, m - k0);
    __syncthreads();
    stage_floats(pts, src + (size_t)3 * k0, 3 * count, threadIdx.x, blockDim.x);
    __syncthreads();
    if (has) {
      for (int k = 0; k < count; ++k) {
        const double d = (double)sqdist(ux, uy, uz, pts[3 * k], pts[3 * k + 1], pts[3 * k + 2]);
        if (d < best1) {
          best3 = best2; besti3 = besti2;
          best2 = best1; besti2 = besti1;
          best1 = d; besti1 = k0 + k;
        } else if (d < best2) {
          best3 = best2; besti3 = besti2;
          best2 = d; besti2 = k0 + k;
        } else if (d < best3) {
          best3 = d; besti3 = k0 + k;
        }
      }
    }
  }
  if (has) {
    float *od = dist2 + ((size_t)b * n + j) * 3;
    int *oi = idx + ((size_t)b * n + j) * 3;
    od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;
    oi[0] = besti1; oi[1] = besti2; oi[2] = besti3;
  }
}

// three_interpolate: out[b,c,j] = fma(p3,w3, fma(p1,w1, p2*w2))  (PTX-verified contraction)
__global__ void __launch_bounds__(256)
three_interpolate_kernel(const float *__restrict__ points, const int *__restrict__ idx,
                         const float *__restrict__ weight, int c, int m, int n, float *__restrict__ out) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int *ii = idx + ((size_t)b * n + j) * 3;
  const float *w = weight + ((size_t)b * n + j) * 3;
  const int i1 = ii[0], i2 = ii[1], i3 = ii[2];
  const float w1 = w[0], w2 = w[1], w3 = w[2];
  for (int l = 0; l < c; ++l) {
    const float *p = points + ((size_t)b * c + l) * m;
    out[((size_t)b * c + l) * n + j] =
        __fmaf_rn(__ldg(p + i3), w3, __fmaf_rn(__ldg(p + i1), w1, __fmul_rn(__ldg(p + i2), w2)));
  }
}

__global__ void __launch_bounds__(256)
three_interpolate_grad_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx,
                              const float *__restrict__ weight, int c, int n, int m,
                              float *__restrict__ grad_points) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int *ii = idx + ((size_t)b * n + j) * 3;
  const float *w = weight + ((size_t)b * n + j) * 3;
  const int i1 = ii[0], i2 = ii[1], i3 = ii[2];
  const float w1 = w[0], w2 = w[1], w3 = w[2];
  for (int l = 0; l < c; ++l) {
    const float g = __ldg(grad_out + ((size_t)b * c + l) * n + j);
    float *gp = grad_points + ((size_t)b * c + l) * m;
    atomicAdd(gp + i1, __fmul_rn(g, w1));
    atomicAdd(gp + i2, __fmul_rn(g, w2));
    atomicAdd(gp + i3, __fmul_rn(g, w3));
  }
}

// ------------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------------
struct FpsGeom {
  int BS, lgBS, Qmax, lgT, spt;
};

FpsGeom fps_geom(int N, int T_max_lg) {
  FpsGeom g;
  g.BS = sv::ref_opt_n_threads(N);
  g.lgBS = 0;
  while ((1 << g.lgBS) < g.BS) ++g.lgBS;
  g.Qmax = (N + g.BS - 1) / g.BS;
  g.lgT = g.lgBS < T_max_lg ? g.lgBS : T_max_lg;
  g.spt = (g.BS >> g.lgT) * g.Qmax;
  return g;
}

template <int SPT, bool FUSE>
int launch_fps_warp(const float *xyz, int B, int N, int m, const FpsGeom &g, int *idx, float *new_xyz, float radius,
                    int nsample, int *ball_idx, cudaStream_t st) {
  size_t smem = (size_t)((3 * N + 3) & ~3) * 4 + 32 * 4;
  if (FUSE) smem += 32 * 4 + (size_t)32 * (nsample | 1) * 4;
  auto kern = fps_warp_kernel<SPT, FUSE>;
  if (smem > 48 * 1024) {
    int rc = sv::cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (rc) return rc;
  }
  kern<<<B, 32, smem, st>>>(xyz, N, m, g.BS, g.lgBS, g.Qmax, g.lgT, g.spt, idx, new_xyz, radius, nsample, ball_idx);
  return sv::after_launch();
}

template <bool FUSE>
int dispatch_fps_warp(const float *xyz, int B, int N, int m, int *idx, float *new_xyz, float radius, int nsample,
                      int *ball_idx, cudaStream_t st) {
  const FpsGeom g = fps_geom(N, 5);
#define SV_CASE(S) \
  if (g.spt <= S) return launch_fps_warp<S, FUSE>(xyz, B, N, m, g, idx, new_xyz, radius, nsample, ball_idx, st)
  SV_CASE(1);
  SV_CASE(2);
  SV_CASE(4);
  SV_CASE(8);
  SV_CASE(16);
  SV_CASE(32);
#undef SV_CASE
  return SV_ERR_INVALID_ARG;
}

template <int SPT>
int launch_fps_block(const float *xyz, int B, int N, int m, const FpsGeom &g, int *idx, float *new_xyz,
                     cudaStream_t st) {
  const size_t smem = (size_t)((3 * N + 3) & ~3) * 4 + SPT * 4 + 2 * 8 * 2 * 4;
  auto kern = fps_block_kernel<SPT>;
  int rc = sv::cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (rc) return rc;
  kern<<<B, 256, smem, st>>>(xyz, N, m, g.Qmax, g.spt, idx, new_xyz);
  return sv::after_launch();
}

// scratch for the generic FPS path (one row of N floats per cloud), grown on demand per device
struct Scratch {
  float *ptr = nullptr;
  size_t bytes = 0;
};
Scratch g_scratch[64];

int get_scratch(size_t bytes, cudaStream_t st, float **out) {
  int dev = 0;
  int rc = sv::cuda_status(cudaGetDevice(&dev));
  if (rc) return rc;
  if (dev < 0 || dev >= 64) return SV_ERR_INVALID_ARG;
  Scratch &s = g_scratch[dev];
  if (s.bytes < bytes) {
    if (s.ptr) {
      rc = sv::cuda_status(cudaStreamSynchronize(st));
      if (rc) return rc;
      cudaFree(s.ptr);
      s.ptr = nullptr;
      s.bytes = 0;
    }
    rc = sv::cuda_status(cudaMalloc(&s.ptr, bytes));
    if (rc) return rc;
    s.bytes = bytes;
  }
  *out = s.ptr;
  return SV_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int sv_version(void) { return 100; }

const char *sv_status_string(int status) {
  switch (status) {
    case SV_OK: return "ok";
    case SV_ERR_INVALID_ARG: return "invalid argument";
    case SV_ERR_CUDA: return "CUDA error";
    default: return "unknown status";
  }
}

int sv_last_cuda_error(void) { return sv::t_last_cuda_error; }
const char *sv_last_cuda_error_string(void) { return cudaGetErrorString((cudaError_t)sv::t_last_cuda_error); }
unsigned long long sv_launch_count(void) { return sv::g_launches.load(); }

int sv_fps_f32(const float *xyz, int B, int N, int m, int *idx, float *new_xyz, void *stream) {
  if (B < 0 || N < 1 || m < 0) return SV_ERR_INVALID_ARG;
  if (B == 0 || m == 0) return SV_OK;
  if (!xyz || !idx) return SV_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (N <= 1024) return dispatch_fps_warp<false>(xyz, B, N, m, idx, new_xyz, 0.f, 0, nullptr, st);
  if (N <= 8192) {
    const FpsGeom g = fps_geom(N, 8);  // BS = 512, T = 256 -> spt = 2*Qmax
    if (g.spt <= 8) return launch_fps_block<8>(xyz, B, N, m, g, idx, new_xyz, st);
    if (g.spt <= 16) return launch_fps_block<16>(xyz, B, N, m, g, idx, new_xyz, st);
    if (g.spt <= 32) return launch_fps_block<32>(xyz, B, N, m, g, idx, new_xyz, st);
  }
  {
    // large clouds: register-resident slices on co-resident CTAs (csrc/fps_coop.cu); shapes that do not fit fall through
    const int rc_coop = sv::fps_coop(xyz, B, N, m, idx, new_xyz, st);
    if (rc_coop != SV_ERR_INVALID_ARG) return rc_coop;
  }
  const FpsGeom g = fps_geom(N, 5);
  float *temp = nullptr;
  int rc = get_scratch((size_t)B * N * sizeof(float), st, &temp);
  if (rc) return rc;
  fps_generic_kernel<<<B, 1024, 0, st>>>(xyz, N, m, g.BS, g.lgBS, g.Qmax, temp, idx, new_xyz);
  return sv::after_launch();
}

int sv_fps_ballquery_f32(const float *xyz, int B, int N, int m, float radius, int nsample, int *fps_idx,
                         float *new_xyz, int *ball_idx, void *stream) {
  if (B < 0 || N < 1 || N > 1024 || m < 0 || nsample < 1 || nsample > 256) return SV_ERR_INVALID_ARG;
  if (B == 0 || m == 0) return SV_OK;
  if (!xyz || !fps_idx || !ball_idx) return SV_ERR_INVALID_ARG;
  return dispatch_fps_warp<true>(xyz, B, N, m, fps_idx, new_xyz, radius, nsample, ball_idx, (cudaStream_t)stream);
}

int sv_ball_query_f32(const float *new_xyz, const float *xyz, int B, int N, int M, float radius, int nsample,
                      int *idx, void *stream) {
  if (B < 0 || N < 0 || M < 0 || nsample < 0) return SV_ERR_INVALID_ARG;
  if (B == 0 || M == 0 || nsample == 0) return SV_OK;
  if (!new_xyz || !xyz || !idx) return SV_ERR_INVALID_ARG;
  if (nsample > 512) return SV_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int W = (M + 31) / 32;
  if (W > 8) W = 8;
  // keep the staging rows within the 227 KB budget
  while (W > 1 && (size_t)3 * BQ_TILE * 4 + (size_t)W * 32 * 4 + (size_t)W * 32 * (nsample | 1) * 4 > 200 * 1024) W >>= 1;
  const size_t smem = (size_t)3 * BQ_TILE * 4 + (size_t)W * 32 * 4 + (size_t)W * 32 * (nsample | 1) * 4;
  if (smem > 48 * 1024) {
    int rc = sv::cuda_status(
        cudaFuncSetAttribute(ball_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (rc) return rc;
  }
  const long long nblk = (long long)((M + 32 * W - 1) / (32 * W)) * B;
  if (nblk > 0x7fffffffLL) return SV_ERR_INVALID_ARG;
  ball_query_kernel<<<(unsigned)nblk, 32 * W, smem, st>>>(new_xyz, xyz, N, M, radius, nsample, idx);
  return sv::after_launch();
}

int sv_group_points_f32(const float *points, const int *idx, int B, int C, int N, int NP, int NS, float *out,
                        void *stream) {
  if (B < 0 || C < 0 || N < 0 || NP < 0 || NS < 0) return SV_ERR_INVALID_ARG;
  const long long E = (long long)NP * NS;
  if (B == 0 || C == 0 || E == 0) return SV_OK;
  if (!points || !idx || !out || E > 0x7fffffffLL) return SV_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (E % 4 == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  // the batch rides on grid.y (<= 65535): larger batches go out in chunks (the reference kernels have no batch limit)
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    const float *pp = points + (size_t)b0 * C * N;
    const int *ii = idx + (size_t)b0 * E;
    float *oo = out + (size_t)b0 * C * E;
    if (vec) {
      dim3 grid((unsigned)((E / 4 + 255) / 256), nb);
      group_points_kernel<true><<<grid, 256, 0, st>>>(pp, ii, C, N, (int)E, oo);
    } else {
      dim3 grid((unsigned)((E + 255) / 256), nb);
      group_points_kernel<false><<<grid, 256, 0, st>>>(pp, ii, C, N, (int)E, oo);
    }
    const int rc = sv::after_launch();
    if (rc) return rc;
  }
  return SV_OK;
}

int sv_group_points_grad_f32(const float *grad_out, const int *idx, int B, int C, int N, int NP, int NS,
                             float *grad_points, void *stream) {
  if (B < 0 || C < 0 || N < 0 || NP < 0 || NS < 0) return SV_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long long E = (long long)NP * NS;
  if ((long long)B * C * N > 0) {
    if (!grad_points) return SV_ERR_INVALID_ARG;
    int rc = sv::cuda_status(cudaMemsetAsync(grad_points, 0, (size_t)B * C * N * sizeof(float), st));
    if (rc) return rc;
  }
  if (B == 0 || C == 0 || E == 0) return SV_OK;
  if (!grad_out || !idx || E > 0x7fffffffLL) return SV_ERR_INVALID_ARG;
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    dim3 grid((unsigned)((E + 255) / 256), nb);
    group_points_grad_kernel<<<grid, 256, 0, st>>>(grad_out + (size_t)b0 * C * E, idx + (size_t)b0 * E, C, N, (int)E,
                                                   grad_points + (size_t)b0 * C * N);
    const int rc = sv::after_launch();
    if (rc) return rc;
  }
  return SV_OK;
}

int sv_gather_points_f32(const float *points, const int *idx, int B, int C, int N, int M, float *out, void *stream) {
  return sv_group_points_f32(points, idx, B, C, N, M, 1, out, stream);
}

int sv_gather_points_grad_f32(const float *grad_out, const int *idx, int B, int C, int N, int M, float *grad_points,
                              void *stream) {
  return sv_group_points_grad_f32(grad_out, idx, B, C, N, M, 1, grad_points, stream);
}

int sv_three_nn_f32(const float *unknown, const float *known, int B, int n, int m, float *dist2, int *idx,
                    void *stream) {
  if (B < 0 || n < 0 || m < 0) return SV_ERR_INVALID_ARG;
  if (B == 0 || n == 0) return SV_OK;
  if (!unknown || (!known && m > 0) || !dist2 || !idx) return SV_ERR_INVALID_ARG;
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    dim3 grid((n + 255) / 256, nb);
    three_nn_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(unknown + (size_t)b0 * n * 3, known ? known + (size_t)b0 * m * 3 : known, n,
                                                            m, dist2 + (size_t)b0 * n * 3, idx + (size_t)b0 * n * 3);
    const int rc = sv::after_launch();
    if (rc) return rc;
  }
  return SV_OK;
}

int sv_three_interpolate_f32(const float *points, const int *idx, const float *weight, int B, int c, int m, int n,
                             float *out, void *stream) {
  if (B < 0 || c < 0 || m < 0 || n < 0) return SV_ERR_INVALID_ARG;
  if (B == 0 || c == 0 || n == 0) return SV_OK;
  if (!points || !idx || !weight || !out) return SV_ERR_INVALID_ARG;
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    dim3 grid((n + 255) / 256, nb);
    three_interpolate_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(points + (size_t)b0 * c * m, idx + (size_t)b0 * n * 3,
                                                                     weight + (size_t)b0 * n * 3, c, m, n, out + (size_t)b0 * c * n);
    const int rc = sv::after_launch();
    if (rc) return rc;
  }
  return SV_OK;
}

int sv_three_interpolate_grad_f32(const float *grad_out, const int *idx, const float *weight, int B, int c, int n,
                                  int m, float *grad_points, void *stream) {
  if (B < 0 || c < 0 || m < 0 || n < 0) return SV_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if ((long long)B * c * m > 0) {
    if (!grad_points) return SV_ERR_INVALID_ARG;
    int rc = sv::cuda_status(cudaMemsetAsync(grad_points, 0, (size_t)B * c * m * sizeof(float), st));
    if (rc) return rc;
  }
  if (B == 0 || c == 0 || n == 0) return SV_OK;
  if (!grad_out || !idx || !weight) return SV_ERR_INVALID_ARG;
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    dim3 grid((n + 255) / 256, nb);
    three_interpolate_grad_kernel<<<grid, 256, 0, st>>>(grad_out + (size_t)b0 * c * n, idx + (size_t)b0 * n * 3,
                                                        weight + (size_t)b0 * n * 3, c, n, m, grad_points + (size_t)b0 * c * m);
    const int rc = sv::after_launch();
    if (rc) return rc;
  }
  return SV_OK;
}

}  // extern "C"
