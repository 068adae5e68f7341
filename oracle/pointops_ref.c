/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the product path
 * (sceneverse_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it.
 *
 * CPU restatement of the reference's PointNet++ CUDA operators.  Paths cited below are relative
 * to /root/reference/modules/third_party/pointnet2/_ext_src/ .
 * The reference has NO CPU implementation of these ops (src/sampling.cpp:34,61,83 assert
 * "CPU not supported"), so this file restates the *.cu kernels thread-for-thread:
 * each "CUDA thread" is simulated, shared-memory trees are executed literally, and every
 * floating-point expression uses the contraction order nvcc emits for the reference source
 * (verified from PTX generated from the reference .cu with nvcc 12.9):
 *     d2 = fma(dz,dz, fma(dx,dx, dy*dy))
 * Compile with -ffp-contract=off so that gcc does not add or remove contractions.
 *
 * Parity pinning: the reference ships no golden vectors for these ops; this
 * restatement is pinned against the reference's own CUDA sources compiled for sm_100a
 * (oracle/_ref, built by oracle/build_ref_ext.py) run on the B200 box, with the outputs
 * committed under tests/golden/ (tests/golden/README.md says which files).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TOTAL_THREADS 512 /* include/cuda_utils.h:11 */

/* include/cuda_utils.h:13-19 — opt_n_threads: 2^floor(log(n)/log(2)) clamped to [1,512]. */
int svref_opt_n_threads(int work_size) {
  const int pow_2 = (int)(log((double)work_size) / log(2.0));
  int v = 1 << pow_2;
  if (v > TOTAL_THREADS) v = TOTAL_THREADS;
  if (v < 1) v = 1;
  return v;
}

/* include/cuda_utils.h:21-28 — opt_block_config(x,y) -> (x_threads, y_threads). */
void svref_opt_block_config(int x, int y, int *xt, int *yt) {
  const int x_threads = svref_opt_n_threads(x);
  int y_threads = svref_opt_n_threads(y);
  if (y_threads > TOTAL_THREADS / x_threads) y_threads = TOTAL_THREADS / x_threads;
  if (y_threads < 1) y_threads = 1;
  *xt = x_threads;
  *yt = y_threads;
}

static inline float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
  /* (a-b) per component; nvcc contraction: fma(dz,dz, fma(dx,dx, dy*dy)) */
  const float dx = ax - bx, dy = ay - by, dz = az - bz;
  return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* src/sampling_gpu.cu:59-65 — __update */
static inline void fps_update(float *dists, int *dists_i, int idx1, int idx2) {
  const float v1 = dists[idx1], v2 = dists[idx2];
  const int i1 = dists_i[idx1], i2 = dists_i[idx2];
  dists[idx1] = fmaxf(v1, v2);
  dists_i[idx1] = v2 > v1 ? i2 : i1;
}

/*
 * src/sampling_gpu.cu:69-173 furthest_point_sampling_kernel<block_size>, launched by
 * :175-229 with block_size = opt_n_threads(n), grid = b; host init src/sampling.cpp:70-76
 * (idxs zeros, temp = 1e10).
 * xyz (B,N,3) f32 -> idx (B,m) i32.
 */
void svref_furthest_point_sampling(const float *xyz, int B, int N, int m, int *idx) {
  if (B <= 0) return;
  memset(idx, 0, sizeof(int) * (size_t)B * (size_t)(m > 0 ? m : 0));
  if (m <= 0) return;
  const int block_size = svref_opt_n_threads(N);
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    const float *dataset = xyz + (size_t)b * N * 3;
    int *idxs = idx + (size_t)b * m;
    float *temp = (float *)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
    float dists[TOTAL_THREADS];
    int dists_i[TOTAL_THREADS];
    for (int k = 0; k < N; ++k) temp[k] = 1e10f;
    int old = 0;
    idxs[0] = old;
    for (int j = 1; j < m; ++j) {
      const float x1 = dataset[old * 3 + 0], y1 = dataset[old * 3 + 1], z1 = dataset[old * 3 + 2];
      for (int tid = 0; tid < block_size; ++tid) {
        int besti = 0;
        float best = -1.0f;
        for (int k = tid; k < N; k += block_size) {
          const float x2 = dataset[k * 3 + 0], y2 = dataset[k * 3 + 1], z2 = dataset[k * 3 + 2];
          const float mag = fmaf(z2, z2, fmaf(x2, x2, y2 * y2));
          if ((double)mag <= 1e-3) continue; /* double compare, sampling_gpu.cu:100-101 */
          const float d = sqdist(x2, y2, z2, x1, y1, z1);
          const float d2 = fminf(d, temp[k]);
          temp[k] = d2;
          besti = d2 > best ? k : besti;
          best = d2 > best ? d2 : best;
        }
        dists[tid] = best;
        dists_i[tid] = besti;
      }
      /* shared-memory tree, sampling_gpu.cu:115-168 */
      for (int s = block_size / 2; s >= 1; s >>= 1)
        for (int tid = 0; tid < s; ++tid) fps_update(dists, dists_i, tid, tid + s);
      old = dists_i[0];
      idxs[j] = old;
    }
    free(temp);
  }
}

/* src/sampling_gpu.cu:8-20 gather_points_kernel: points (B,C,N), idx (B,M) -> out (B,C,M). */
void svref_gather_points(const float *points, const int *idx, int B, int C, int N, int M,
                         float *out) {
#pragma omp parallel for collapse(2)
  for (int i = 0; i < B; ++i)
    for (int l = 0; l < C; ++l)
      for (int j = 0; j < M; ++j) {
        const int a = idx[(size_t)i * M + j];
        out[((size_t)i * C + l) * M + j] = points[((size_t)i * C + l) * N + a];
      }
}

/* src/sampling_gpu.cu:34-47 gather_points_grad_kernel (atomicAdd; here a serial sum in j order —
 * compare with a tolerance, the reference's own order is non-deterministic). */
void svref_gather_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int M,
                              float *grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)B * C * N);
#pragma omp parallel for collapse(2)
  for (int i = 0; i < B; ++i)
    for (int l = 0; l < C; ++l)
      for (int j = 0; j < M; ++j) {
        const int a = idx[(size_t)i * M + j];
        grad_points[((size_t)i * C + l) * N + a] += grad_out[((size_t)i * C + l) * M + j];
      }
}

/*
 * src/ball_query_gpu.cu:9-44 query_ball_point_kernel; host zero-init src/ball_query.cpp:19-21.
 * new_xyz (B,M,3), xyz (B,N,3) -> idx (B,M,nsample).
 */
void svref_ball_query(const float *new_xyz, const float *xyz, int B, int N, int M, float radius,
                      int nsample, int *idx) {
  memset(idx, 0, sizeof(int) * (size_t)B * M * nsample);
  const float radius2 = radius * radius; /* fp32 product, ball_query_gpu.cu:22 */
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    const float *p = xyz + (size_t)b * N * 3;
    const float *q = new_xyz + (size_t)b * M * 3;
    int *o = idx + (size_t)b * M * nsample;
    for (int j = 0; j < M; ++j) {
      const float nx = q[j * 3 + 0], ny = q[j * 3 + 1], nz = q[j * 3 + 2];
      for (int k = 0, cnt = 0; k < N && cnt < nsample; ++k) {
        const float d2 = sqdist(nx, ny, nz, p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
        if (d2 < radius2) {
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) o[j * nsample + l] = k;
          o[j * nsample + cnt] = k;
          ++cnt;
        }
      }
    }
  }
}

/* src/group_points_gpu.cu:8-28: points (B,C,N), idx (B,NP,NS) -> out (B,C,NP,NS). */
void svref_group_points(const float *points, const int *idx, int B, int C, int N, int NP, int NS,
                        float *out) {
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int l = 0; l < C; ++l)
      for (int j = 0; j < NP; ++j)
        for (int k = 0; k < NS; ++k) {
          const int ii = idx[((size_t)b * NP + j) * NS + k];
          out[(((size_t)b * C + l) * NP + j) * NS + k] = points[((size_t)b * C + l) * N + ii];
        }
}

/* src/group_points_gpu.cu:43-64 (atomicAdd scatter; serial (j,k) order here). */
void svref_group_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int NP,
                             int NS, float *grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)B * C * N);
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int l = 0; l < C; ++l)
      for (int j = 0; j < NP; ++j)
        for (int k = 0; k < NS; ++k) {
          const int ii = idx[((size_t)b * NP + j) * NS + k];
          grad_points[((size_t)b * C + l) * N + ii] +=
              grad_out[(((size_t)b * C + l) * NP + j) * NS + k];
        }
}

/*
 * src/interpolate_gpu.cu:9-59 three_nn_kernel: unknown (B,n,3), known (B,m,3) ->
 * dist2 (B,n,3) f32, idx (B,n,3) i32.  Best distances are kept in double (init 1e40), the
 * candidate distance is an fp32 value promoted for the strict '<' compares, results are
 * rounded back to fp32 on store (1e40 -> +inf when m < 3).
 */
void svref_three_nn(const float *unknown, const float *known, int B, int n, int m, float *dist2,
                    int *idx) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    const float *u = unknown + (size_t)b * n * 3;
    const float *kn = known + (size_t)b * m * 3;
    for (int j = 0; j < n; ++j) {
      const float ux = u[j * 3 + 0], uy = u[j * 3 + 1], uz = u[j * 3 + 2];
      double best1 = 1e40, best2 = 1e40, best3 = 1e40;
      int besti1 = 0, besti2 = 0, besti3 = 0;
      for (int k = 0; k < m; ++k) {
        const float d = sqdist(ux, uy, uz, kn[k * 3 + 0], kn[k * 3 + 1], kn[k * 3 + 2]);
        if (d < best1) {
          best3 = best2; besti3 = besti2;
          best2 = best1; besti2 = besti1;
          best1 = d; besti1 = k;
        } else if (d < best2) {
          best3 = best2; besti3 = besti2;
          best2 = d; besti2 = k;
        } else if (d < best3) {
          best3 = d; besti3 = k;
        }
      }
      float *od = dist2 + ((size_t)b * n + j) * 3;
      int *oi = idx + ((size_t)b * n + j) * 3;
      od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;
      oi[0] = besti1; oi[1] = besti2; oi[2] = besti3;
    }
  }
}

/* src/interpolate_gpu.cu:72-101: points (B,c,m), idx/weight (B,n,3) -> out (B,c,n);
 * nvcc contraction (PTX-verified): fma(p3,w3, fma(p1,w1, p2*w2)). */
void svref_three_interpolate(const float *points, const int *idx, const float *weight, int B,
                             int c, int m, int n, float *out) {
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < n; ++j) {
        const float *w = weight + ((size_t)b * n + j) * 3;
        const int *ii = idx + ((size_t)b * n + j) * 3;
        const float *p = points + ((size_t)b * c + l) * m;
        out[((size_t)b * c + l) * n + j] =
            fmaf(p[ii[2]], w[2], fmaf(p[ii[0]], w[0], p[ii[1]] * w[1]));
      }
}

/* src/interpolate_gpu.cu:116-143 (atomicAdd scatter of grad*w; serial order here). */
void svref_three_interpolate_grad(const float *grad_out, const int *idx, const float *weight,
                                  int B, int c, int n, int m, float *grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)B * c * m);
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < n; ++j) {
        const float *w = weight + ((size_t)b * n + j) * 3;
        const int *ii = idx + ((size_t)b * n + j) * 3;
        const float g = grad_out[((size_t)b * c + l) * n + j];
        float *gp = grad_points + ((size_t)b * c + l) * m;
        gp[ii[0]] += g * w[0];
        gp[ii[1]] += g * w[1];
        gp[ii[2]] += g * w[2];
      }
}
