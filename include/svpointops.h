/*
 * libsvpointops — C-ABI of the B200 (sm_100a) PointNet++ point operators.
 *
 * This is the drop-in boundary for the reference's pybind extension `pointnet2._ext`
 * (reference: modules/third_party/pointnet2/_ext_src/src/bindings.cpp:6-19).  Every entry point
 * takes raw DEVICE pointers + sizes + a cudaStream_t (passed as void*), enqueues on that stream,
 * does not synchronise, does not allocate, and returns an int status (0 = OK).  Output buffers are
 * fully overwritten (no pre-zeroing needed; the *_grad entry points zero their output themselves,
 * matching the reference's torch::zeros allocation).  All tensors are contiguous fp32 / int32, the
 * same preconditions the reference asserts (include/utils.h:5-25).
 *
 * Index results (fps / ball_query / three_nn idx) are bit-exact with the reference kernels;
 * copies (gather / group) are exact; *_grad use fp32 atomics like the reference (order not fixed).
 */
#ifndef SVPOINTOPS_H
#define SVPOINTOPS_H

#ifdef __cplusplus
extern "C" {
#endif

#define SV_OK 0
#define SV_ERR_INVALID_ARG 1   /* null pointer, negative size, unsupported shape */
#define SV_ERR_CUDA 2          /* launch failed; see sv_last_cuda_error() */

/* library identification */
int sv_version(void);
const char *sv_status_string(int status);
/* cudaError_t of the last failing launch in this thread (0 if none), and its string */
int sv_last_cuda_error(void);
const char *sv_last_cuda_error_string(void);
/* number of kernel launches this library has enqueued so far in this process (bench bookkeeping) */
unsigned long long sv_launch_count(void);

/* replaces furthest_point_sampling (src/sampling.cpp:66-87, kernel src/sampling_gpu.cu:69-173).
 * xyz (B,N,3) -> idx (B,m) i32.  new_xyz may be NULL; when given receives xyz[idx] as (B,m,3)
 * (fuses the gather_points call that always follows, pointnet2_modules.py:54-58). */
int sv_fps_f32(const float *xyz, int B, int N, int m, int *idx, float *new_xyz, void *stream);

/* replaces gather_points (src/sampling.cpp:15-38, src/sampling_gpu.cu:8-20).
 * points (B,C,N), idx (B,M) -> out (B,C,M) */
int sv_gather_points_f32(const float *points, const int *idx, int B, int C, int N, int M,
                         float *out, void *stream);

/* replaces gather_points_grad (src/sampling.cpp:40-65, src/sampling_gpu.cu:34-47).
 * grad_out (B,C,M), idx (B,M) -> grad_points (B,C,N) (zeroed here, then atomic scatter-add) */
int sv_gather_points_grad_f32(const float *grad_out, const int *idx, int B, int C, int N, int M,
                              float *grad_points, void *stream);

/* replaces ball_query (src/ball_query.cpp:8-32, src/ball_query_gpu.cu:9-44).
 * new_xyz (B,M,3), xyz (B,N,3) -> idx (B,M,nsample) i32 */
int sv_ball_query_f32(const float *new_xyz, const float *xyz, int B, int N, int M, float radius,
                      int nsample, int *idx, void *stream);

/* replaces group_points (src/group_points.cpp:12-36, src/group_points_gpu.cu:8-28).
 * points (B,C,N), idx (B,NP,NS) -> out (B,C,NP,NS) */
int sv_group_points_f32(const float *points, const int *idx, int B, int C, int N, int NP, int NS,
                        float *out, void *stream);

/* replaces group_points_grad (src/group_points.cpp:38-62, src/group_points_gpu.cu:43-64).
 * grad_out (B,C,NP,NS), idx (B,NP,NS) -> grad_points (B,C,N) (zeroed here) */
int sv_group_points_grad_f32(const float *grad_out, const int *idx, int B, int C, int N, int NP,
                             int NS, float *grad_points, void *stream);

/* replaces three_nn (src/interpolate.cpp:14-40, src/interpolate_gpu.cu:9-59).
 * unknown (B,n,3), known (B,m,3) -> dist2 (B,n,3) f32, idx (B,n,3) i32 */
int sv_three_nn_f32(const float *unknown, const float *known, int B, int n, int m, float *dist2,
                    int *idx, void *stream);

/* replaces three_interpolate (src/interpolate.cpp:42-70, src/interpolate_gpu.cu:72-101).
 * points (B,c,m), idx (B,n,3), weight (B,n,3) -> out (B,c,n) */
int sv_three_interpolate_f32(const float *points, const int *idx, const float *weight, int B, int c,
                             int m, int n, float *out, void *stream);

/* replaces three_interpolate_grad (src/interpolate.cpp:71-99, src/interpolate_gpu.cu:116-143).
 * grad_out (B,c,n), idx, weight (B,n,3) -> grad_points (B,c,m) (zeroed here) */
int sv_three_interpolate_grad_f32(const float *grad_out, const int *idx, const float *weight, int B,
                                  int c, int n, int m, float *grad_points, void *stream);

/* Fused sample+query used by the new set-abstraction path (one pass over xyz): equals
 * sv_fps_f32(xyz -> fps_idx,new_xyz) followed by sv_ball_query_f32(new_xyz, xyz -> ball_idx).
 * fps_idx (B,m) i32, new_xyz (B,m,3) f32, ball_idx (B,m,nsample) i32. Requires N <= 1024. */
int sv_fps_ballquery_f32(const float *xyz, int B, int N, int m, float radius, int nsample,
                         int *fps_idx, float *new_xyz, int *ball_idx, void *stream);

/* Two-level fused sampling for a PointNet++ set-abstraction stack, one pass over xyz (32 <= N <= 1024):
 * level 1 = sv_fps_ballquery_f32(xyz, m, radius, nsample); level 2 (m2 > 0, requires m == 32) = the
 * same on the (B,32,3) centres of level 1 with (m2, radius_2, nsample2).  Replaces the reference's
 * furthest_point_sample + gather_operation + ball_query of two consecutive PointnetSAModules
 * (pointnet2_modules.py:54-58, pointnet2_utils.py:331); the ball-query distances are taken from the
 * FPS sweep (bit-identical).  Outputs as in sv_fps_ballquery_f32, per level. */
int sv_sa_sample_f32(const float *xyz, int B, int N, int m, float radius, int nsample, int *fps_idx,
                     float *new_xyz, int *ball_idx, int m2, float radius_2, int nsample2, int *fps_idx2,
                     float *new_xyz2, int *ball_idx2, void *stream);

/* ---- device-side input pipeline -----------------------------------------------------------------------
 * Builds the model's per-object tensors from ragged raw scene data on the GPU: replaces the per-object numpy loop of
 * data/datasets/base.py:697-741 (`_obj_processing_post`) and the padding of data/datasets/dataset_wrapper.py:62-72.
 * raw_points (total,6) f32 [xyz rgb]; slot_offsets (n_slots+1) int64 CSR over the B*O object slots (empty range = padded
 * slot).  Per slot: obj_locs (6) = [mean xyz, max-min] of the raw points; P points sampled uniformly (without replacement
 * when the object has >= P points, with replacement otherwise), centred on the SAMPLE mean, divided by the sample's max norm
 * (1 when < 1e-6), colours untouched; empty slots: all-ones points, zero locs, mask 0.  sample_idx (n_slots,P) i32 may be
 * NULL (receives the chosen raw indices, -1 for empty slots).  Randomness = counter hash of (seed, slot, k). */
int sv_scene_prep_f32(const float *raw_points, const long long *slot_offsets, int n_slots, int P, unsigned long long seed,
                      float *obj_fts, float *obj_locs, unsigned char *obj_masks, int *sample_idx, void *stream);
/* BERT masked-LM corruption (data/data_utils.py:76-104 `random_word`): for tokens with attention_mask != 0, with probability
 * mask_ratio the label is the original id and the token becomes [MASK] (80 %), a uniform random id (10 %) or stays (10 %);
 * every other label is -1.  ids / attention_mask / out_ids / labels: (n_tokens) int64. */
int sv_token_mask(const long long *ids, const long long *attention_mask, long long n_tokens, float mask_ratio,
                  long long mask_token_id, long long vocab_size, unsigned long long seed, long long *out_ids,
                  long long *labels, void *stream);
/* random_point_cloud (data/data_utils.py:107-121): out[i] = valid[i] && u_i >= drop_ratio (bytes). */
int sv_coin_mask(const unsigned char *valid, long long n, float drop_ratio, unsigned long long seed, unsigned char *out,
                 void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SVPOINTOPS_H */
