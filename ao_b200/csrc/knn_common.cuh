// knn_common.cuh — pieces shared by the TILE and GRID kNN kernels.
#pragma once
#include "common.cuh"

namespace aopt {

// Squared distance with the reference's exact rounding sequence (see oracle/knn_oracle.c):
// FADD x3, FMUL (y term), FFMA (x term), FFMA (z term).  Explicit intrinsics so that the result
// does not depend on the compiler's contraction choices.
__device__ __forceinline__ float dist2_ref(float qx, float qy, float qz, float x, float y, float z) {
    float dx = __fsub_rn(qx, x), dy = __fsub_rn(qy, y), dz = __fsub_rn(qz, z);
    float t = __fmul_rn(dy, dy);
    t = __fmaf_rn(dx, dx, t);
    return __fmaf_rn(dz, dz, t);
}

// Scene of point q: first s with q < offset[s]  (== the linear scan of knn_query_cuda_kernel.cu:45-56).
// Returns b when q is past the last offset.
__device__ __forceinline__ int find_segment(int q, const int *__restrict__ offset, int b) {
    int lo = 0, hi = b;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (q < __ldg(offset + mid)) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

// k best candidates as a sorted list held in registers (all indices compile-time).
// LEX = false: candidates are offered in ascending index, so "strictly smaller than the k-th" plus a
//              stable insert already yields ascending (d2, idx).
// LEX = true : candidates arrive in arbitrary order (grid traversal); compare (d2, idx) pairs.
template <int K, bool LEX>
struct TopK {
    float d[K];
    int id[K];

    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < K; ++i) { d[i] = 1e10f; id[i] = -1; }
    }
    __device__ __forceinline__ float worst() const { return d[K - 1]; }

    __device__ __forceinline__ static bool before(float da, int ia, float db, int ib) {
        if (LEX) return da < db || (da == db && ia < ib);
        return da < db;
    }
    __device__ __forceinline__ void offer(float d2, int i) {
        bool acc = LEX ? (d2 < d[K - 1] || (d2 == d[K - 1] && i < id[K - 1] && d2 < 1e10f)) : (d2 < d[K - 1]);
        if (acc) {
            d[K - 1] = d2;
            id[K - 1] = i;
#pragma unroll
            for (int j = K - 1; j > 0; --j) {
                bool sw = before(d[j], id[j], d[j - 1], id[j - 1]);
                float td = d[j]; int ti = id[j];
                d[j] = sw ? d[j - 1] : td;   id[j] = sw ? id[j - 1] : ti;
                d[j - 1] = sw ? td : d[j - 1]; id[j - 1] = sw ? ti : id[j - 1];
            }
        }
    }
    __device__ __forceinline__ void store(int *idx_row, float *d2_row, int nsample) const {
#pragma unroll
        for (int i = 0; i < K; ++i) {
            if (i < nsample) { idx_row[i] = id[i]; d2_row[i] = d[i]; }
        }
    }
};

}  // namespace aopt
