/*
 * ao_pointops.h — C ABI of libao_pointops.so, the B200 (sm_100a) point-operator library that
 * replaces the native side of the reference's `pointops._C` for the PTv2m2 hot path.
 *
 * Boundary conventions (mirror the reference launchers):
 *   - every entry point takes raw DEVICE pointers and sizes, borrows them for the call only,
 *     allocates nothing and keeps no state (re-entrant; safe from the autograd thread);
 *   - outputs and scratch ("workspace") are caller-allocated; ask *_workspace_bytes() first;
 *   - kernels are launched on the caller's stream (`stream` is a cudaStream_t passed as void*;
 *     NULL = legacy default stream, which is what the reference launchers use implicitly);
 *   - the return value is an aopt status (0 = AOPT_OK).  The reference launchers return void and
 *     never check errors; here argument errors and launch errors are reported.
 *   - offsets are cumulative END indices (`offset[b]` = one past the last point of scene b), int32,
 *     exactly the reference's offset-encoded batch layout (libs/pointops/functions/query.py:22).
 *
 * Reference interfaces replaced (paths relative to /root/reference/libs/pointops/src):
 *   aopt_knn_query               knn_query/knn_query_cuda_kernel.h:15   knn_query_cuda_launcher
 *   aopt_grouping_forward        grouping/grouping_cuda_kernel.h:14     grouping_forward_cuda_launcher
 *   aopt_grouping_backward       grouping/grouping_cuda_kernel.h:15     grouping_backward_cuda_launcher
 *   aopt_interpolation_forward   interpolation/interpolation_cuda_kernel.h  interpolation_forward_cuda_launcher
 *   aopt_interpolation_backward  interpolation/interpolation_cuda_kernel.h  interpolation_backward_cuda_launcher
 *   aopt_farthest_point_sampling sampling/sampling_cuda_kernel.h        farthest_point_sampling_cuda_launcher
 *   aopt_aggregation_forward/backward   aggregation/aggregation_cuda_kernel.h   (PTv1 "share-planes" layout)
 *   aopt_subtraction_forward/backward   subtraction/subtraction_cuda_kernel.h
 * and the torch / third-party op chains of the PTv2m2 caller
 * (pointcept/models/point_transformer_v2/point_transformer_v2m2_base.py):
 *   aopt_group_xyz, aopt_gather_sub_*     :109,:112 + libs/pointops/functions/grouping.py:36-60
 *   aopt_gva_*                            :119-128   (softmax over k, mask, grouped weighted sum)
 *   aopt_voxel_*, aopt_pool_*             :244-269   (voxel_grid + segment_csr mean/max)
 *   aopt_interp_weights                   libs/pointops/functions/interpolation.py:15-17
 *   aopt_grid_sample_keys, aopt_voxel_pick, aopt_sphere_dist2, aopt_select_rows
 *                                         pointcept/datasets/transform.py:792-896,968-979 (GridSample, SphereCrop)
 *   aopt_vote_accumulate                  pointcept/engines/test.py:106-113 (tester: softmax + fragment vote)
 *   aopt_csr_build                        (new) transpose of the neighbour graph → atomic-free backward
 */
#ifndef AO_POINTOPS_H_
#define AO_POINTOPS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *aopt_stream_t; /* cudaStream_t */

enum {
    AOPT_OK = 0,
    AOPT_ERR_INVALID_ARGUMENT = 1, /* bad size / NULL pointer / k out of range */
    AOPT_ERR_WORKSPACE = 2,        /* workspace missing or too small */
    AOPT_ERR_LAUNCH = 3,           /* cudaGetLastError() != cudaSuccess after a launch */
    AOPT_ERR_UNSUPPORTED = 4
};

#define AOPT_MAX_NSAMPLE 128 /* reference limit: knn_query_cuda_kernel.cu:82-83 */

/* kNN search strategy (all three are exact and return identical results). */
enum { AOPT_KNN_AUTO = 0, AOPT_KNN_TILE = 1, AOPT_KNN_GRID = 2 };
/* OR-ed into `method`: write sqrt(dist2) (IEEE round-to-nearest, == torch.sqrt of the squared output) into the
 * distance buffer — what the reference's Python returns (functions/query.py:24) without a second kernel. */
enum { AOPT_KNN_SQRT_DIST = 0x100 };

const char *aopt_version(void);
const char *aopt_status_string(int status);
/* cudaGetErrorString of the last launch error seen by this thread (for AOPT_ERR_LAUNCH). */
const char *aopt_last_cuda_error(void);
/* Cumulative number of kernels this library has enqueued in the process (all threads); bench.py
 * reports the difference over its timed region as "gpu_launches". */
unsigned long long aopt_kernel_launches(void);
/* Tuning switches for A/B measurements and tests (never needed for correctness; every setting gives the same
 * results bit for bit): "csr_impl" 1 = radix sort / 2 = count-fill-rank, "gva_bwd" 1 = fused / 2 = two kernels,
 * "voxel_sort" 1 = compact keys, 3 passes / 2 = wide keys, 6 passes, "knn_sample" 1 = cell edge from the bounding box /
 * 2 = sampled, "pdl" 1 = programmatic dependent launch inside the small-kernel chains / 2 = off, "knn_topk" 1 = top-k of the
 * GRID query kernel as a shared-memory heap / 2 = as a sorted list in registers, "knn_pend" 1 = accepted candidates
 * go through a per-lane pending list drained every eight candidates / 2 = inserted in place, "l2pf" 1 = the GVA kernels
 * request the next work item's peb block into L2 with a bulk prefetch (the default) / 2 = off, "knn_site" 1 = GRID query kernel with a
 * single scan / insert site (knn_grid1_kernel) / 2 = the default kernel; 0 = library default.
 * Initial values come from AOPT_CSR_IMPL / AOPT_GVA_BWD / AOPT_VOXEL_SORT / AOPT_KNN_SAMPLE / AOPT_PDL / AOPT_L2PF / AOPT_KNN_SITE / AOPT_KNN_TOPK /
 * AOPT_KNN_PEND. */
int aopt_set_tuning(const char *name, int value);

/* ---- offset-encoded batch layout ---------------------------------------------------------- */
/* batch[i] = scene of point i (int64, like pointcept/models/utils.py:11-24). */
int aopt_offset2batch(int n, int b, const int *offset, int64_t *batch, aopt_stream_t stream);

/* ---- kNN ---------------------------------------------------------------------------------- */
/* xyz (n,3), new_xyz (m,3) row-major fp32; offset/new_offset (b,) int32.
 * idx (m,nsample) int32, -1 padded; dist2 (m,nsample) fp32 SQUARED distances, 1e10 padded, rows
 * ascending by (dist2, idx).  dist2 = fma(dz,dz,fma(dx,dx,dy*dy)), d = query - candidate: the
 * bit pattern the reference kernel produces when built for sm_100a. */
size_t aopt_knn_workspace_bytes(int n, int m, int b, int nsample, int method);
int aopt_knn_query(int m, int nsample, int n, int b, const float *xyz, const float *new_xyz,
                   const int *offset, const int *new_offset, int *idx, float *dist2, int method,
                   void *workspace, size_t workspace_bytes, aopt_stream_t stream);
/* The same search over several INDEPENDENT point sets at once: the sets are concatenated scene by scene into one
 * offset-encoded batch (the search never crosses a scene), and index_base (b,) int32 — may be NULL — is subtracted
 * from the valid indices of every scene, so that each set gets indices relative to its own first point.  One call
 * (one grid build, one query launch) instead of one per set: the pyramid levels 1..L of a PTv2m2 forward
 * (...v2m2_base.py:223 once per BlockSequence; functions/interpolation.py:14 once per unpooling stage). */
int aopt_knn_query_multi(int m, int nsample, int n, int b, const float *xyz, const float *new_xyz,
                         const int *offset, const int *new_offset, const int *index_base, int *idx,
                         float *dist2, int method, void *workspace, size_t workspace_bytes,
                         aopt_stream_t stream);

/* ---- farthest point sampling (PTv1 caller: point_transformer_seg.py:101) --------------------- */
/* Arguments of farthest_point_sampling_cuda_launcher (sampling/sampling_cuda_kernel.h): b scenes,
 * n_max = size of the largest scene, xyz (n,3), offset / new_offset (b) cumulative ends, tmp (n)
 * running squared distances pre-filled with 1e10 by the caller (functions/sampling.py:19), idx
 * (new_offset[b-1]) output: idx[new_offset[s-1]] = first point of scene s, then repeatedly the point
 * farthest from everything chosen so far.  Same results as the reference kernel for every input,
 * equal-distance ties included (fps.cu header).  One thread-block cluster per scene. */
int aopt_farthest_point_sampling(int b, int n_max, const float *xyz, const int *offset,
                                 const int *new_offset, float *tmp, int *idx, aopt_stream_t stream);

/* ---- per-sample data transforms (pointcept/datasets/transform.py) ------------- */
/* GridSample front half (transform.py:806-811): cell = floor(coord / grid) per axis (f64 != 0: fp64
 * division as NumPy >= 2 evaluates `fp32 array / np.array(float)`; 0: fp32 as NumPy 1.x did), minus the
 * per-axis minimum; keys = hash(cell) ^ 2^63 (hash_type 0 = FNV64-1A :881-896, 1 = ravel :864-878) so that
 * an int64 sort orders like the reference's uint64 argsort.  stats[0..2] = per-axis minimum of
 * floor(coord / grid) (min_coord = stats * grid, :808), stats[3..5] = maximum. */
int aopt_grid_sample_keys(int n, const float *coord, double grid_x, double grid_y, double grid_z,
                          int f64, int hash_type, int *cell, int64_t *keys, int *stats,
                          aopt_stream_t stream);
/* pick[v] = order[idx_ptr[v] + r[v] % count[v]] (transform.py:813-817; r == NULL: r_const for every
 * voxel = test mode part number, :841-843).  order / idx_ptr from aopt_voxel_partition. */
int aopt_voxel_pick(int n_vox, const int *idx_ptr, const int *order, const int64_t *r,
                    long long r_const, int64_t *pick, aopt_stream_t stream);
/* SphereCrop (transform.py:973-975): dist2[i] = sum(square(coord[i] - centre)) in fp32, numpy's order. */
int aopt_sphere_dist2(int n, const float *coord, float cx, float cy, float cz, float *dist2,
                      aopt_stream_t stream);
/* out[i, :] = src[index[i], :] for rows of `words_per_row` 4-byte words (data_dict[key][idx], any dtype). */
int aopt_select_rows(long long rows, int words_per_row, const void *src, const int64_t *index,
                     void *out, aopt_stream_t stream);

/* ---- transposed neighbour graph (CSR) ------------------------------------------------------ */
/* idx: n_entries int32 values in [-1, n_src) (flattened (m,nsample)).  Produces rowptr (n_src+1)
 * and perm (n_entries): perm[rowptr[j] .. rowptr[j+1]) = flat positions p with idx[p] == j in
 * ASCENDING p (deterministic summation order).  negative_mode 0: idx<0 entries are dropped;
 * 1: idx<0 is wrapped to idx+n_src (python negative indexing, interpolation.py:21). */
size_t aopt_csr_workspace_bytes(int n_src, int64_t n_entries);
int aopt_csr_build(int n_src, int64_t n_entries, const int *idx, int negative_mode, int *rowptr,
                   int *perm, void *workspace, size_t workspace_bytes, aopt_stream_t stream);

/* ---- grouping (neighbour gather) ----------------------------------------------------------- */
/* output[(p)*out_stride + ch] = input[idx[p]*c + ch], zeros when idx[p] < 0; p in [0, m*nsample). */
int aopt_grouping_forward(int m, int nsample, int c, const float *input, const int *idx,
                          float *output, int out_stride, aopt_stream_t stream);
/* grad_input[j,:] = scale * sum over CSR row j of grad_output[perm[e]*go_stride + :]  (no atomics). */
int aopt_grouping_backward(int n, int c, const float *grad_output, int go_stride, const int *rowptr,
                           const int *perm, float scale, float *grad_input, aopt_stream_t stream);
/* Backward of out[j,s,:] = key[idx[j,s],:] - query[j,:] when queries and sources are the same n points
 * (GroupedVectorAttention, …v2m2_base.py:109,112): grad_key[j] = sum over CSR row j of grad rows,
 * grad_query[j] = -sum_s grad[j,s,:], in ONE pass over grad (n,nsample,c): the second read of every row
 * is served by L2.  Same summation orders as aopt_grouping_backward / aopt_sum_over_k. */
int aopt_relation_backward(int n, int nsample, int c, const float *grad, const int *rowptr,
                           const int *perm, float *grad_key, float *grad_query, aopt_stream_t stream);
/* out[p*out_stride + 0..2] = (xyz[idx[p]] - new_xyz[p / nsample]) * sign(idx[p]+1). */
int aopt_group_xyz(int m, int nsample, const float *xyz, const float *new_xyz, const int *idx,
                   float *out, int out_stride, aopt_stream_t stream);
/* out[m,s,:] = key[idx[m,s],:] - query[m,:]   (key row = 0 when idx < 0). */
int aopt_gather_sub_forward(int m, int nsample, int c, const float *key, const float *query,
                            const int *idx, float *out, aopt_stream_t stream);
/* out[m,:] = scale * sum_s grad[m,s,:]   (grad_query of gather_sub uses scale = -1). */
int aopt_sum_over_k(int m, int nsample, int c, const float *grad, float scale, float *out,
                    aopt_stream_t stream);

/* ---- GroupedVectorAttention softmax-over-k weighted aggregation ---------------------------- */
/* value (n_src,c) un-gathered; peb (n,nsample,c) or NULL; logits (n,nsample,g); idx (n,nsample).
 * out[n, gi*I+i] = sum_s (value[idx[n,s], gi*I+i] + peb[n,s,gi*I+i]) * softmax_s(logits[n,:,gi])[s]
 *                  * sign(idx[n,s]+1),   I = c/g.     prob (n,nsample,g) receives the UNMASKED
 * softmax (saved for backward) when not NULL. */
int aopt_gva_forward(int n, int nsample, int c, int g, const float *value, const float *peb,
                     const float *logits, const int *idx, float *out, float *prob,
                     aopt_stream_t stream);
/* grad_peb (n,nsample,c) (may be NULL) and grad_logits (n,nsample,g). */
int aopt_gva_backward_query(int n, int nsample, int c, int g, const float *grad_out,
                            const float *value, const float *peb, const float *prob, const int *idx,
                            float *grad_peb, float *grad_logits, aopt_stream_t stream);
/* grad_value (n_src,c) through the CSR of idx (no atomics). */
int aopt_gva_backward_value(int n_src, int nsample, int c, int g, const float *grad_out,
                            const float *prob, const int *rowptr, const int *perm,
                            float *grad_value, aopt_stream_t stream);
/* Both of the above in ONE kernel for self-attention (queries == sources, n_src == n): the CSR walk of
 * a thread's source row runs in the shadow of the DRAM latency of its query item.  Same results bit for
 * bit; falls back to the two kernels for layouts outside the specialised path.  Autograd of
 * ...v2m2_base.py:110,119-128 in one pass (SURVEY.md 8d "Fused GVA bwd"). */
int aopt_gva_backward(int n, int nsample, int c, int g, const float *grad_out, const float *value,
                      const float *peb, const float *prob, const int *idx, const int *rowptr,
                      const int *perm, float *grad_peb, float *grad_logits, float *grad_value,
                      aopt_stream_t stream);

/* ---- GridPool ----------------------------------------------------------------------------- */
/* start (b,3) = per-scene minimum corner (segment_csr(coord, ptr, "min")). */
int aopt_segment_min3(int n, int b, const float *coord, const int *offset, float *start,
                      aopt_stream_t stream);
/* keys[i] = order-preserving packing of (scene, z, y, x) voxel coordinates with
 * cell_d = (int)((coord_d - start_d) / grid_size) in fp32 — same cells and same sort order as
 * torch_cluster.grid_cluster's Σ cell_d·stride_d.  status_flag (device int, may be NULL) is set
 * to 1 if a cell coordinate does not fit the packing (18 bits per axis, 10 bits scene). */
int aopt_voxel_keys(int n, int b, const float *coord, const int *offset, const float *start,
                    float grid_size, int64_t *keys, int *status_flag, aopt_stream_t stream);
/* Voxel partition from the keys sorted ascending (sorted_keys, order64 = the stable argsort, both
 * (n) int64): order32 (n) = point ids by voxel, cluster32/cluster64 (n) = voxel id of every point
 * (voxels numbered in ascending key order = torch.unique(sorted=True), …v2m2_base.py:260-262),
 * idx_ptr (n+1 allocated, n_vox+1 used), new_offset (b) int64 = cumulative voxel count per scene
 * (:267-268), meta[0] = n_vox.  No host synchronisation inside. */
size_t aopt_voxel_partition_workspace_bytes(int n);
int aopt_voxel_partition(int n, int b, const int64_t *sorted_keys, const int64_t *order64,
                         const int *offset, int *order32, int *cluster32, int64_t *cluster64,
                         int *idx_ptr, int64_t *new_offset, int *meta, void *workspace,
                         size_t workspace_bytes, aopt_stream_t stream);
/* The whole front half of GridPool (...v2m2_base.py:246-264) in one call, all on the device: per-scene
 * bounding boxes, voxel keys packed into the fewest bits that hold (scene, z, y, x), a stable radix sort of
 * the points by key (ceil(bits / 11) passes; `max_passes` in [1,6] are enqueued and the unneeded ones return
 * at once), voxel boundaries, ids, idx_ptr and per-scene voxel offsets.  start (b,3) may be NULL = per-scene
 * minimum (segment_csr(coord, ptr, "min")).  Outputs as aopt_voxel_partition.  meta (8 ints): [0] number of
 * voxels, [1] flags (1: a point below `start` or a key wider than 64 bits; 2: more than max_passes passes
 * needed - call again with 6), [2] passes needed, [3] key bits. */
size_t aopt_voxel_grid_workspace_bytes(int n, int b);
int aopt_voxel_grid(int n, int b, const float *coord, const int *offset, const float *start,
                    float grid_size, int max_passes, int *order32, int *cluster32, int64_t *cluster64,
                    int *idx_ptr, int64_t *new_offset, int *meta, void *workspace,
                    size_t workspace_bytes, aopt_stream_t stream);
/* Points sorted by voxel: order (n) = point ids, idx_ptr (n_vox+1).  out_feat/argmax (n_vox,c):
 * max over the voxel and the ORIGINAL id of the first maximal point; out_coord (n_vox,3) = mean
 * (sequential sum in `order`, divided by the count). */
int aopt_pool_forward(int n_vox, int c, const float *feat, const float *coord, const int *order,
                      const int *idx_ptr, float *out_feat, int *argmax, float *out_coord,
                      aopt_stream_t stream);
/* grad_feat[i,ch] = grad_out[cluster[i],ch] if argmax[cluster[i],ch] == i else 0. */
int aopt_pool_backward(int n, int c, const float *grad_out, const int *argmax, const int *cluster,
                       float *grad_feat, aopt_stream_t stream);

/* ---- tester fragment vote ------------------------------------------------------------------ */
/* pred[index[r], :] += softmax(logits[r, :]) for r in [0, rows): the per-fragment accumulation of
 * pointcept/engines/test.py:106-113 (F.softmax, then `pred[idx_part[bs:be], :] += pred_part[bs:be]`) in one
 * kernel.  logits (rows, c) fp32, index (rows) int64 with DISTINCT values inside one call (one fragment; negative
 * values wrap like a python index), pred (n_pred, c) fp32 updated in place.  *bad_flag (optional, device int) is
 * set to 1 if an index falls outside [-n_pred, n_pred); such rows are skipped. */
int aopt_vote_accumulate(int rows, int c, long long n_pred, const float *logits, const long long *index,
                         float *pred, int *bad_flag, aopt_stream_t stream);

/* ---- three-NN inverse-distance interpolation ---------------------------------------------- */
/* weight[n,i] = r_i / sum_j r_j,  r = 1/(sqrt(dist2)+1e-8)  (interpolation.py:15-17). */
int aopt_interp_weights(int n, int k, const float *dist2, float *weight, aopt_stream_t stream);
/* output[n,:] = sum_i input[wrap(idx[n,i]),:] * weight[n,i];  wrap(j) = j<0 ? j+m : j. */
int aopt_interpolation_forward(int n, int c, int k, int m, const float *input, const int *idx,
                               const float *weight, float *output, aopt_stream_t stream);
/* grad_input[j,:] = sum over CSR row j of grad_output[perm[e]/k,:] * weight[perm[e]]. */
int aopt_interpolation_backward(int m, int c, int k, const float *grad_output, const float *weight,
                                const int *rowptr, const int *perm, float *grad_input,
                                aopt_stream_t stream);

/* ---- fused positional-bias MLP ("next" row) ------------------------------------------ */
/* peb = Linear(C,C)(ReLU(BatchNorm(Linear(3,C)(pos)))) on all rows = N*nsample neighbour rows, the
 * `linear_p_bias` branch of GroupedVectorAttention (…v2m2_base.py:88-93,116-118), without materialising
 * the hidden (rows,C) tensors: training-mode BatchNorm statistics follow in closed form from the mean and
 * covariance of pos (aopt_pos_moments).  Widths: aopt_pe_mlp_supported(c) (48 and 96); bf16 tensor-core
 * products with fp32 accumulation. */
int aopt_pe_mlp_supported(int c);
/* moments: 9 doubles = Σp (3), Σ xx xy xz yy yz zz (6) over the rows of pos (rows,3). */
size_t aopt_pos_moments_workspace_bytes(void);
int aopt_pos_moments(int64_t rows, const float *pos, double *moments, void *workspace,
                     size_t workspace_bytes, aopt_stream_t stream);
/* `state` (aopt_pe_mlp_state_bytes(c), caller-allocated) carries the folded BatchNorm maps and the bf16
 * copies of W2 from forward to backward.  use_batch_stats = 1: training (statistics from `moments`);
 * 0: evaluation (running_mean / running_var). */
size_t aopt_pe_mlp_state_bytes(int c);
/* Optional auxiliary head: aux_w (ga,c), ga <= 16, aux_out (rows,ga) = aux_w · h (no bias) — the hidden
 * activation h is shared, so a Linear applied to peb folds into the same kernel as one more MMA column tile
 * (ptv2 uses it for weight_encoding[0]∘linear_p_bias[3], which removes the (rows,c) relation tensor). */
int aopt_pe_mlp_forward(int64_t rows, int c, const float *pos, const double *moments, const float *w1,
                        const float *b1, const float *gamma, const float *beta, const float *running_mean,
                        const float *running_var, float eps, int use_batch_stats, const float *w2,
                        const float *b2, float *out, const float *aux_w, int ga, float *aux_out, void *state,
                        size_t state_bytes, aopt_stream_t stream);
/* stats_out (3c floats) = batch mean | biased variance | rstd of the first layer (running-stat update). */
int aopt_pe_mlp_stats(int c, const void *state, float *stats_out, aopt_stream_t stream);
/* Parameter gradients for grad (rows,c) = dL/dpeb.  One pass over grad; deterministic (no atomics). */
size_t aopt_pe_mlp_backward_workspace_bytes(int64_t rows, int c);
int aopt_pe_mlp_backward(int64_t rows, int c, const float *pos, const double *moments, const float *w1,
                         const float *gamma, int use_batch_stats, const float *grad, const void *state,
                         float *grad_w1, float *grad_b1, float *grad_gamma, float *grad_beta, float *grad_w2,
                         float *grad_b2, int ga, const float *grad_aux, float *grad_aux_w, void *workspace,
                         size_t workspace_bytes, aopt_stream_t stream);

/* ---- BatchNorm-shaped element-wise work around the dense layers ("next" row) ----------- */
/* Element types of the (rows, c) activations below. */
#define AOPT_F32 0
#define AOPT_BF16 1
/* Scratch for the calls below: `width` = 2*c (aopt_bn_act_*) or 3*g + g*g (aopt_we_tail_*). */
size_t aopt_dense_workspace_bytes(int width);
/* out = [ReLU]( [residual +] [row_scale[row] *] BatchNorm_train(x) ) over the rows of x (rows, c): PointBatchNorm on
 * (N, C) / (N*k, C) tensors with the nn.ReLU, DropPath scale and residual add that follow it in a PTv2 block
 * (point_transformer_v2m2_base.py:25-45,187-197; Linear -> PointBatchNorm -> ReLU triples :86-93,240-242,288-295).
 * Batch statistics in fp32 with fp64 combination; x and out/residual are fp32 or bf16 (x_dtype / out_dtype; residual has
 * out's type); c % 4 == 0, c <= 1024 (aopt_bn_act_supported).  ldx = row stride of x in elements (>= c, a multiple of 4):
 * x may be a column block of a wider matrix, e.g. the q or k part of a fused q|k|v GEMM output; out is dense.  stats_out (2c floats) = batch mean | rstd, kept for the
 * backward pass; running_mean / running_var (optional) are updated like nn.BatchNorm1d does (momentum, unbiased
 * variance).  mean_shift (optional, c floats) is added to the batch mean in the running-mean update only: a Linear bias
 * in front of a training-mode BatchNorm cancels in the output, so the caller may leave it out of x.  residual, row_scale
 * may be NULL; relu = 0 / 1.  batches_tracked (optional): device int64 counter incremented by one
 * (nn.BatchNorm1d.num_batches_tracked). */
int aopt_bn_act_supported(int c);
int aopt_bn_act_forward(int64_t rows, int c, const void *x, int64_t ldx, int x_dtype, const float *gamma, const float *beta, float eps,
                        const void *residual, const float *row_scale, int relu, void *out, int out_dtype,
                        float *stats_out, float *running_mean, float *running_var, float momentum,
                        const float *mean_shift, long long *batches_tracked, void *workspace, size_t workspace_bytes,
                        aopt_stream_t stream);
/* Backward of the above.  out = the forward result (only read when relu was 1: pass NULL otherwise).  grad_x has x's
 * type and row stride ldgx; grad_residual (optional, out's type) = grad_out masked by the ReLU.  Deterministic (no atomics). */
int aopt_bn_act_backward(int64_t rows, int c, const void *grad_out, const void *out, int out_dtype, const void *x,
                         int64_t ldx, int x_dtype, const float *stats, const float *gamma, const float *row_scale,
                         void *grad_x, int64_t ldgx, void *grad_residual, float *grad_gamma, float *grad_beta, void *workspace,
                         size_t workspace_bytes, aopt_stream_t stream);
/* Tail of GroupedVectorAttention.weight_encoding (point_transformer_v2m2_base.py:94-99,120) on the (rows = N*nsample, g)
 * tensors:  u = rel + upe + cst  (upe (rows, g) and cst (g) optional),  logits = W2 * ReLU(BatchNorm_train(u)) + b2.
 * g in {6, 12} (aopt_we_tail_supported); w2 (g, g) row-major [out][in]; b2 optional; stats_out (2g) = mean | rstd.
 * rel == NULL selects gather mode: rel[row] = kp[idx[row]] - qp[row / nsample] (kp, qp (N, g); idx (N, nsample), key row = 0
 * where idx < 0; rows = N * nsample) is formed while loading — the g-wide aopt_gather_sub_forward folded in. */
int aopt_we_tail_supported(int g);
int aopt_we_tail_forward(int64_t rows, int g, const float *rel, const float *kp, const float *qp, const int *idx,
                         int nsample, const float *upe, const float *cst, const float *gamma,
                         const float *beta, float eps, const float *w2, const float *b2, float *logits, float *stats_out,
                         float *running_mean, float *running_var, float momentum, long long *batches_tracked,
                         void *workspace, size_t workspace_bytes, aopt_stream_t stream);
/* grad_u (rows, g) is the gradient of rel and of upe alike (the gradient of cst is identically zero: it sits in
 * front of a training-mode BatchNorm).  grad_w2 (g, g), grad_b2 / grad_gamma / grad_beta (g).  Deterministic. */
int aopt_we_tail_backward(int64_t rows, int g, const float *rel, const float *kp, const float *qp, const int *idx,
                          int nsample, const float *upe, const float *cst, const float *grad_logits, const float *stats, const float *gamma, const float *beta,
                          const float *w2, float *grad_u, float *grad_gamma, float *grad_beta, float *grad_b2,
                          float *grad_w2, void *workspace, size_t workspace_bytes, aopt_stream_t stream);

/* Small dense helpers of the Linear layers around the operators above.
 * aopt_col_sum: out (c floats) = column sums of x (rows, c), row stride ldx — a bias gradient; workspace
 *   aopt_dense_workspace_bytes(2*c).
 * aopt_copy_cols: dst[r, :] = src[r, :] (+ bias) for a c-wide column block, each side with its own row stride and element
 *   type (the v block of a fused q|k|v product <-> a dense fp32 tensor).
 * aopt_skinny_linear / aopt_skinny_dgrad / aopt_skinny_wgrad: a Linear with a handful (g) of outputs over hundreds of
 *   thousands of rows — weight_encoding[0] applied to key / query (g = groups), the segmentation head (g = num_classes) — and,
 *   with the roles of input and output exchanged, the patch-embedding projection (g = in_channels); w (g, c) fp32,
 *   the first layer of linear_p_bias (g = 3 coordinates); g in {3, 4, 6, 9, 12, 13, 19, 20} (aopt_skinny_wgrad_supported):
 *   out (rows, g) fp32 = x wᵀ (+ bias);  grad_x (rows, c) = grad (rows, g) w;  grad_w (g, c) fp32 = gradᵀ x (workspace
 *   aopt_dense_workspace_bytes(g*c); deterministic). */
int aopt_col_sum(int64_t rows, int c, const void *x, int64_t ldx, int x_dtype, float *out, void *workspace,
                 size_t workspace_bytes, aopt_stream_t stream);
int aopt_copy_cols(int64_t rows, int c, const void *src, int64_t ld_src, int src_dtype, const float *bias, void *dst,
                   int64_t ld_dst, int dst_dtype, aopt_stream_t stream);
int aopt_skinny_wgrad_supported(int g, int c);
int aopt_skinny_linear(int64_t rows, int g, int c, const void *x, int64_t ldx, int x_dtype, const float *w,
                       const float *bias, float *out, aopt_stream_t stream);
int aopt_skinny_dgrad(int64_t rows, int g, int c, const float *grad, const float *w, void *grad_x, int64_t ldgx, int x_dtype,
                      aopt_stream_t stream);
int aopt_skinny_wgrad(int64_t rows, int g, int c, const void *grad, int grad_dtype, const void *x, int64_t ldx, int x_dtype,
                      float *out, void *workspace, size_t workspace_bytes, aopt_stream_t stream);

/* ---- PTv1-layout fused ops kept for API parity --------------------------------------------- */
/* output[n,ch] = sum_s (input[idx[n,s],ch] + position[n,s,ch]) * weight[n,s,ch % w_c]. */
int aopt_aggregation_forward(int n, int nsample, int c, int w_c, const float *input,
                             const float *position, const float *weight, const int *idx,
                             float *output, aopt_stream_t stream);
/* grad_position (n,nsample,c), grad_weight (n,nsample,w_c) per query; grad_input through the CSR. */
int aopt_aggregation_backward(int n, int nsample, int c, int w_c, const float *input,
                              const float *position, const float *weight, const int *idx,
                              const int *rowptr, const int *perm, const float *grad_output,
                              float *grad_input, float *grad_position, float *grad_weight,
                              aopt_stream_t stream);
/* output[n,s,:] = input1[n,:] - input2[idx[n,s],:].  Backward: grad_input1 = aopt_sum_over_k(+1),
 * grad_input2 = aopt_grouping_backward(scale = -1). */
int aopt_subtraction_forward(int n, int nsample, int c, const float *input1, const float *input2,
                             const int *idx, float *output, aopt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AO_POINTOPS_H_ */
