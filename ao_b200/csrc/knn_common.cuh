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
// LEX = true : candidates arrive in arbitrary order (grid traversal); compare (d2, idx) pairs.  d2 >= +0
//              (a sum of squares, never -0 or negative), so its bit pattern orders like the value and a pair is
//              one 64-bit key (d2 bits : idx) — two integer compares instead of three float/int ones.
//
// Insert = K independent "is the candidate before entry j" tests, then every entry picks its new value from
// {left neighbour, candidate, itself}.  The previous bubble-from-the-end form made each step wait for the swap
// before it (a ~60-deep dependent chain for K = 16); whenever ANY lane of a warp accepts a candidate the whole
// warp walks the insert, so on the small levels (one warp per scheduler) the chain latency was the kernel time
// (profiles/r01i: knn_grid 120 us for 12.5k and for 50k queries alike).  Same result: the list is sorted and the
// order is strict, so "shift everything after the insertion point" equals the bubble.
#ifndef AOPT_TOPK_HD
#define AOPT_TOPK_HD __device__ __forceinline__
#define AOPT_F2U(x) __float_as_uint(x)
#endif
template <int K, bool LEX>
struct TopK {
    float d[K];
    int id[K];

    AOPT_TOPK_HD void init() {
#pragma unroll
        for (int i = 0; i < K; ++i) { d[i] = 1e10f; id[i] = -1; }
    }
    AOPT_TOPK_HD float worst() const { return d[K - 1]; }

    AOPT_TOPK_HD static unsigned long long key(float dv, int iv) {
        return ((unsigned long long)AOPT_F2U(dv) << 32) | (unsigned)iv;
    }
    AOPT_TOPK_HD static bool before(float da, int ia, float db, int ib) {
        if (LEX) return key(da, ia) < key(db, ib);
        return da < db;
    }
    AOPT_TOPK_HD void offer(float d2, int i) {
        // d2 < 1e10f: the reference never accepts a candidate at or beyond its initial heap value
        // (knn_query_cuda_kernel.cu:85-93); it also keeps NaN distances and the (1e10, -1) padding apart
        const bool acc = LEX ? (d2 < 1e10f && before(d2, i, d[K - 1], id[K - 1])) : (d2 < d[K - 1]);
        if (acc) {
            bool lt[K];
#pragma unroll
            for (int j = 0; j < K; ++j) lt[j] = before(d2, i, d[j], id[j]);
#pragma unroll
            for (int j = K - 1; j > 0; --j) {
                d[j] = lt[j - 1] ? d[j - 1] : (lt[j] ? d2 : d[j]);
                id[j] = lt[j - 1] ? id[j - 1] : (lt[j] ? i : id[j]);
            }
            d[0] = lt[0] ? d2 : d[0];
            id[0] = lt[0] ? i : id[0];
        }
    }
    __device__ __forceinline__ void store(int *idx_row, float *d2_row, int nsample, bool root = false) const {
#pragma unroll
        for (int i = 0; i < K; ++i) {
            if (i < nsample) { idx_row[i] = id[i]; d2_row[i] = root ? __fsqrt_rn(d[i]) : d[i]; }
        }
    }
};

}  // namespace aopt
