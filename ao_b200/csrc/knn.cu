// knn.cu — exact batched-offset k nearest neighbours.
//
// Replaces /root/reference/libs/pointops/src/knn_query/knn_query_cuda_kernel.cu:60-104 (one thread
// per query scanning its whole scene from global memory, k-entry heap in a 1 KB local-memory
// stack frame) behind the launcher signature of knn_query_cuda_kernel.h:15.
//
// Contract (bit-exact with the reference build on every tie-free row, see oracle/knn_oracle.c):
//   d2 = fma(dz,dz, fma(dx,dx, dy*dy)),  d = query - candidate        (fp32, IEEE, no fast-math)
//   accept iff d2 < current k-th distance (strict); rows ascending by (d2, idx);
//   unused slots: idx = -1, dist2 = 1e10.
//
// TILE method (this file, part 1): candidates of the query's scene are staged through shared
// memory in tiles of 1024 points (float4 x,y,z,-), every thread owns one query and keeps its k best
// in REGISTERS as a sorted list (template K); a candidate is compared against the k-th distance
// first, four at a time, so the sorted insert (K compare-exchanges) only runs ~k·ln(n/k) times per
// query.  All lanes of a warp read the same shared-memory word (broadcast, conflict-free).
// FP32-ALU bound: ~9 issue slots per (query, candidate) pair.
//
// GRID method (part 2, knn_grid.cu) prunes the candidate set with a uniform grid and is what the
// AUTO policy picks for large scenes; both return identical results.
#include <stdlib.h>

#include "common.cuh"
#include "knn_common.cuh"

namespace aopt {

constexpr int kKnnBlock = 256;
constexpr int kKnnSmallBlock = 64;   // few queries (the coarse levels): 64-thread CTAs so the scan spreads over the chip
constexpr int kKnnTile = 1024;

template <int K, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
knn_tile_kernel(int m, int n, int b, int nsample, const float *__restrict__ xyz,
                const float *__restrict__ new_xyz, const int *__restrict__ offset,
                const int *__restrict__ new_offset, int *__restrict__ idx_out,
                float *__restrict__ dist2_out, bool root) {
    __shared__ float4 tile[kKnnTile];
    __shared__ int range_s[2];

    const int q = blockIdx.x * BLOCK + threadIdx.x;
    const bool valid = q < m;
    int start = 0, end = 0;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (valid) {
        int seg = find_segment(q, new_offset, b);
        if (seg < b) {
            start = seg == 0 ? 0 : __ldg(offset + seg - 1);
            end = __ldg(offset + seg);
            if (end > n) end = n;
            if (start < 0) start = 0;
        }
        qx = __ldg(new_xyz + (size_t)q * 3 + 0);
        qy = __ldg(new_xyz + (size_t)q * 3 + 1);
        qz = __ldg(new_xyz + (size_t)q * 3 + 2);
    }
    // Candidate range of the block = union of its queries' scene ranges (consecutive queries →
    // consecutive scenes); warp-reduce, then one shared atomic per warp.
    if (threadIdx.x == 0) { range_s[0] = 0x7fffffff; range_s[1] = 0; }
    __syncthreads();
    {
        const bool has = end > start;
        int wlo = __reduce_min_sync(0xffffffffu, has ? start : 0x7fffffff);
        int whi = __reduce_max_sync(0xffffffffu, has ? end : 0);
        if ((threadIdx.x & 31) == 0) { atomicMin(&range_s[0], wlo); atomicMax(&range_s[1], whi); }
    }
    __syncthreads();
    const int lo = range_s[0], hi = range_s[1];

    TopK<K, false> top;
    top.init();

    for (int base = lo; base < hi; base += kKnnTile) {
        const int cnt = min(kKnnTile, hi - base);
        __syncthreads();  // previous tile fully consumed
        for (int t = threadIdx.x; t < cnt; t += BLOCK) {
            const float *p = xyz + (size_t)(base + t) * 3;
            tile[t] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
        }
        __syncthreads();
        int j = max(start, base) - base;
        const int jhi = min(end, base + cnt) - base;
        for (; j + 4 <= jhi; j += 4) {
            float4 c0 = tile[j], c1 = tile[j + 1], c2 = tile[j + 2], c3 = tile[j + 3];
            float d0 = dist2_ref(qx, qy, qz, c0.x, c0.y, c0.z);
            float d1 = dist2_ref(qx, qy, qz, c1.x, c1.y, c1.z);
            float d2 = dist2_ref(qx, qy, qz, c2.x, c2.y, c2.z);
            float d3 = dist2_ref(qx, qy, qz, c3.x, c3.y, c3.z);
            float mn = fminf(fminf(d0, d1), fminf(d2, d3));
            if (mn < top.worst()) {
                top.offer(d0, base + j);
                top.offer(d1, base + j + 1);
                top.offer(d2, base + j + 2);
                top.offer(d3, base + j + 3);
            }
        }
        for (; j < jhi; ++j) {
            float4 c0 = tile[j];
            top.offer(dist2_ref(qx, qy, qz, c0.x, c0.y, c0.z), base + j);
        }
    }
    if (valid) top.store(idx_out + (size_t)q * nsample, dist2_out + (size_t)q * nsample, nsample, root);
}

// nsample in (32, 128]: same scan, k best kept in a local-memory sorted list (rare path; the
// PTv2m2 configs use k <= 32).
__global__ void __launch_bounds__(kKnnBlock)
knn_tile_bigk_kernel(int m, int n, int b, int nsample, const float *__restrict__ xyz,
                     const float *__restrict__ new_xyz, const int *__restrict__ offset,
                     const int *__restrict__ new_offset, int *__restrict__ idx_out,
                     float *__restrict__ dist2_out, bool root) {
    const int q = blockIdx.x * kKnnBlock + threadIdx.x;
    if (q >= m) return;
    int start = 0, end = 0;
    int seg = find_segment(q, new_offset, b);
    if (seg < b) {
        start = seg == 0 ? 0 : __ldg(offset + seg - 1);
        end = min(__ldg(offset + seg), n);
        if (start < 0) start = 0;
    }
    const float qx = __ldg(new_xyz + (size_t)q * 3 + 0), qy = __ldg(new_xyz + (size_t)q * 3 + 1),
                qz = __ldg(new_xyz + (size_t)q * 3 + 2);
    float bd[AOPT_MAX_NSAMPLE];
    int bi[AOPT_MAX_NSAMPLE];
    for (int i = 0; i < nsample; ++i) { bd[i] = 1e10f; bi[i] = -1; }
    for (int i = start; i < end; ++i) {
        const float *p = xyz + (size_t)i * 3;
        float d2 = dist2_ref(qx, qy, qz, __ldg(p), __ldg(p + 1), __ldg(p + 2));
        if (d2 < bd[nsample - 1]) {
            int j = nsample - 1;
            while (j > 0 && bd[j - 1] > d2) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; --j; }
            bd[j] = d2;
            bi[j] = i;
        }
    }
    for (int i = 0; i < nsample; ++i) {
        idx_out[(size_t)q * nsample + i] = bi[i];
        dist2_out[(size_t)q * nsample + i] = root ? __fsqrt_rn(bd[i]) : bd[i];
    }
}

template <int K>
static void launch_tile(int m, int n, int b, int nsample, const float *xyz, const float *new_xyz,
                        const int *offset, const int *new_offset, int *idx, float *dist2,
                        bool root, cudaStream_t st) {
    // a 256-thread CTA per 256 queries leaves most SMs idle below ~38k queries (level 3: 2868 queries = 12 CTAs)
    if (m <= kKnnSmallBlock * kNumSM * 4)
        knn_tile_kernel<K, kKnnSmallBlock><<<div_up(m, kKnnSmallBlock), kKnnSmallBlock, 0, st>>>(
            m, n, b, nsample, xyz, new_xyz, offset, new_offset, idx, dist2, root);
    else
        knn_tile_kernel<K, kKnnBlock><<<div_up(m, kKnnBlock), kKnnBlock, 0, st>>>(m, n, b, nsample, xyz, new_xyz, offset,
                                                                                  new_offset, idx, dist2, root);
}

int knn_tile_launch(int m, int nsample, int n, int b, const float *xyz, const float *new_xyz,
                    const int *offset, const int *new_offset, int *idx, float *dist2,
                    bool root, cudaStream_t st) {
    if (nsample <= 1) launch_tile<1>(m, n, b, nsample, xyz, new_xyz, offset, new_offset, idx, dist2, root, st);
    else if (nsample <= 4) launch_tile<4>(m, n, b, nsample, xyz, new_xyz, offset, new_offset, idx, dist2, root, st);
    else if (nsample <= 8) launch_tile<8>(m, n, b, nsample, xyz, new_xyz, offset, new_offset, idx, dist2, root, st);
    else if (nsample <= 16) launch_tile<16>(m, n, b, nsample, xyz, new_xyz, offset, new_offset, idx, dist2, root, st);
    else if (nsample <= 32) launch_tile<32>(m, n, b, nsample, xyz, new_xyz, offset, new_offset, idx, dist2, root, st);
    else
        knn_tile_bigk_kernel<<<div_up(m, kKnnBlock), kKnnBlock, 0, st>>>(m, n, b, nsample, xyz, new_xyz, offset,
                                                                        new_offset, idx, dist2, root);
    return check_launch();
}

// part 2 (knn_grid.cu)
size_t knn_grid_workspace_bytes(int n, int m, int b);
int knn_grid_launch(int m, int nsample, int n, int b, const float *xyz, const float *new_xyz,
                    const int *offset, const int *new_offset, int *idx, float *dist2, bool root, void *ws,
                    size_t ws_bytes, cudaStream_t st);

}  // namespace aopt

using namespace aopt;

// Scenes smaller than this (average points per scene) are cheaper to scan than to bin.
// AOPT_KNN_GRID_MIN overrides it (tuning runs).
static int grid_min_points() {
    static int v = [] {
        int d = 2048;
        if (const char *e = getenv("AOPT_KNN_GRID_MIN")) {
            int x = atoi(e);
            if (x > 0) d = x;
        }
        return d;
    }();
    return v;
}

static int pick_method(int n, int m, int b, int nsample, int method) {
    // nothing to search in (no scenes / no candidates): every output row is padding — the tile kernel writes it;
    // the grid needs at least one scene and one point to build (an explicit GRID request is served the same way)
    if (b <= 0 || n <= 0) return AOPT_KNN_TILE;
    if (method == AOPT_KNN_TILE || method == AOPT_KNN_GRID) return method;
    if (nsample > 32) return AOPT_KNN_TILE;
    long long avg = b > 0 ? (long long)n / b : n;
    return avg >= grid_min_points() ? AOPT_KNN_GRID : AOPT_KNN_TILE;
}

extern "C" size_t aopt_knn_workspace_bytes(int n, int m, int b, int nsample, int method) {
    if (n < 0 || m < 0 || b < 0) return 0;
    method &= ~AOPT_KNN_SQRT_DIST;
    if (pick_method(n, m, b, nsample, method) != AOPT_KNN_GRID) return 0;
    return knn_grid_workspace_bytes(n, m, b);
}

// idx[row, :] -= index_base[scene of row] for the valid entries (the -1 padding stays): the second half of
// aopt_knn_query_multi.
__global__ void __launch_bounds__(256)
knn_rebase_kernel(long long total, int nsample, int b, const int *__restrict__ new_offset,
                  const int *__restrict__ index_base, int *__restrict__ idx) {
    aopt::pdl_wait();
    const long long step = (long long)gridDim.x * 256;
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < total; p += step) {
        const int row = (int)(p / nsample);
        const int sc = aopt::find_segment(row, new_offset, b);
        const int v = idx[p];
        if (sc < b && v >= 0) idx[p] = v - __ldg(index_base + sc);
    }
}

extern "C" int aopt_knn_query_multi(int m, int nsample, int n, int b, const float *xyz,
                                    const float *new_xyz, const int *offset, const int *new_offset,
                                    const int *index_base, int *idx, float *dist2, int method, void *workspace,
                                    size_t workspace_bytes, aopt_stream_t stream) {
    if (m < 0 || n < 0 || b < 0 || nsample < 1 || nsample > AOPT_MAX_NSAMPLE) return AOPT_ERR_INVALID_ARGUMENT;
    const bool root = (method & AOPT_KNN_SQRT_DIST) != 0;   // distances instead of squared distances
    method &= ~AOPT_KNN_SQRT_DIST;
    if (method < AOPT_KNN_AUTO || method > AOPT_KNN_GRID) return AOPT_ERR_INVALID_ARGUMENT;
    if (m == 0) return AOPT_OK;
    if (!new_xyz || !idx || !dist2 || (n > 0 && !xyz) || (b > 0 && (!offset || !new_offset)))
        return AOPT_ERR_INVALID_ARGUMENT;
    cudaStream_t st = as_stream(stream);
    int use = pick_method(n, m, b, nsample, method);
    int rc;
    if (use == AOPT_KNN_GRID) {
        if (nsample > 32) return AOPT_ERR_UNSUPPORTED;
        size_t need = knn_grid_workspace_bytes(n, m, b);
        if (!workspace || workspace_bytes < need) return AOPT_ERR_WORKSPACE;
        rc = knn_grid_launch(m, nsample, n, b, xyz, new_xyz, offset, new_offset, idx, dist2, root, workspace,
                             workspace_bytes, st);
    } else {
        rc = knn_tile_launch(m, nsample, n, b, xyz, new_xyz, offset, new_offset, idx, dist2, root, st);
    }
    if (rc != AOPT_OK || !index_base || b == 0) return rc;
    const long long total = (long long)m * nsample;
    knn_rebase_kernel<<<stride_grid(total, 256, 8), 256, 0, st>>>(total, nsample, b, new_offset, index_base, idx);
    return check_launch(1);
}

extern "C" int aopt_knn_query(int m, int nsample, int n, int b, const float *xyz,
                              const float *new_xyz, const int *offset, const int *new_offset,
                              int *idx, float *dist2, int method, void *workspace,
                              size_t workspace_bytes, aopt_stream_t stream) {
    return aopt_knn_query_multi(m, nsample, n, b, xyz, new_xyz, offset, new_offset, nullptr, idx, dist2, method,
                                workspace, workspace_bytes, stream);
}
