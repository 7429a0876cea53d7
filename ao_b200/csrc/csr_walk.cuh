// csr_walk.cuh — batched, software-pipelined walk over the rows of a CSR neighbour graph.
//
// Used by the two WEIGHTED atomic-free backward kernels that reduce over the in-edges of a source row:
//   GVA backward value  grad_value[j] = sum_e prob[perm[e], g] * grad_out[perm[e] / k] (gva.cu)
//   interpolation bwd   grad_in[j]    = sum_e weight[perm[e]] * grad_out[perm[e] / k]  (interp.cu)
// (the unweighted grouping backward keeps its own kernel: see aopt_grouping_backward).
//
// One thread owns (row j, 128-bit channel chunk).  The walk is a dependent chain
// rowptr -> perm -> gathered row.  ncu (profiles/r01h_bv_locality.md) shows the row-at-a-time kernel
// latency-bound: 76 % of the warp samples wait on the long scoreboard while L1, L2, DRAM and the issue
// slots are all < 45 % busy, and Morton-ordering the points lifts the L1 hit rate of the gather from
// 17 % to 64 % (L2 traffic / 2.3) without changing the duration.  Here:
//   * entries are taken B at a time (no remainder loop): the B gathers and the B weights of a batch are
//     requested together.  Slots past the end of the row load entry 0 — unconditional loads, because
//     ptxas serialises predicated ones into three registers — and are dropped by a select;
//   * the perm values of the NEXT batch — of this row or, on its last batch, of the thread's next row —
//     are requested before the current batch is consumed;
//   * rowptr of the row after next is requested one row ahead.
// Measured at level 0 (320k x 16, C=48): gva_backward_value 162 -> 130 us with B = 8.  A cp.async-staged variant
// of the same walk measured the same 125-180 us (it is the ~32 instructions per gathered 16 bytes — 48 % issue
// slots, 67 % L1 — that bound it now), and taught one rule kept below: filler slots must not all read ONE line
// through L2 (entry 0 for every thread via cp.async.cg: 740 us); here they go through L1-allocating loads.
// The grid is sized so that gridDim.x * blockDim.x is a multiple of `chunks`: a thread keeps its channel
// chunk for the whole kernel and only the row advances (no divisions in the loop).
// Entries are accumulated in ascending e (perm is ascending inside a row): the summation order is the
// same fixed order as before — bitwise reproducible.
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace aopt {

constexpr int kWalkBlock = 128;

// AOPT_WALK=old selects the previous row-at-a-time kernels (A/B measurements only).
inline bool use_batched_walk() {
    static const bool on = [] { const char *e = getenv("AOPT_WALK"); return !(e && e[0] == 'o'); }();
    return on;
}
inline int walk_batch() {  // entries per batch: 8 (default) or 4 (AOPT_WALK_B=4)
    static const int b = [] { const char *e = getenv("AOPT_WALK_B"); return (e && e[0] == '4') ? 4 : 8; }();
    return b;
}

// CTAs for a walk over n_rows x chunks items: a multiple of chunks / gcd(chunks, kWalkBlock), at most
// ctas_per_sm x 148, at least one multiple (common.cuh col_grid).
inline int walk_grid(long long n_rows, int chunks, int ctas_per_sm) {
    return col_grid(n_rows, chunks, kWalkBlock, ctas_per_sm);
}

// Policy P:
//   static constexpr bool kWeighted;
//   __device__ float4 load(int p, int ch) const;     the 128-bit piece of the row that entry p refers to
//   __device__ float  weight(int p, int ch) const;   (kWeighted only)
template <int B, class P>
__global__ void __launch_bounds__(kWalkBlock)
csr_walk_kernel(long long n_rows, int chunks, int c_out, const int *__restrict__ rowptr,
                const int *__restrict__ perm, P pol, float scale, float *__restrict__ out) {
    const long long step_items = (long long)gridDim.x * kWalkBlock;  // multiple of chunks (walk_grid)
    const long long t0 = (long long)blockIdx.x * kWalkBlock + threadIdx.x;
    const long long row_step = step_items / chunks;
    long long j = t0 / chunks;
    const int ch = (int)(t0 - j * chunks);
    if (j >= n_rows) return;

    int e = __ldg(rowptr + j), e_end = __ldg(rowptr + j + 1);
    int pn[B];
#pragma unroll
    for (int u = 0; u < B; ++u) pn[u] = (e + u < e_end) ? __ldg(perm + e + u) : 0;
    long long jn = j + row_step;
    int ne = 0, ne_end = 0;
    if (jn < n_rows) { ne = __ldg(rowptr + jn); ne_end = __ldg(rowptr + jn + 1); }

    for (;;) {
        const long long jnn = jn + row_step;
        int nne = 0, nne_end = 0;
        if (jnn < n_rows) { nne = __ldg(rowptr + jnn); nne_end = __ldg(rowptr + jnn + 1); }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (;;) {
            const bool last = e + B >= e_end;
            int p[B];
            float4 v[B];
            float w[B];
#pragma unroll
            for (int u = 0; u < B; ++u) {
                p[u] = pn[u];
                v[u] = pol.load(p[u], ch);  // unconditional: slots past the row end hold p = 0 (a valid entry)
            }
            if (P::kWeighted) {
#pragma unroll
                for (int u = 0; u < B; ++u) w[u] = pol.weight(p[u], ch);
            }
            // next batch: the rest of this row, or the first entries of the thread's next row
            const int pe = last ? ne : e + B, pe_end = last ? ne_end : e_end;
#pragma unroll
            for (int u = 0; u < B; ++u) pn[u] = (pe + u < pe_end) ? __ldg(perm + pe + u) : 0;
            issue_fence();  // all requests of the batch are issued before the first consumer (common.cuh)
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const bool keep = e + u < e_end;
                if (P::kWeighted) {
                    acc.x = keep ? fmaf(v[u].x, w[u], acc.x) : acc.x;
                    acc.y = keep ? fmaf(v[u].y, w[u], acc.y) : acc.y;
                    acc.z = keep ? fmaf(v[u].z, w[u], acc.z) : acc.z;
                    acc.w = keep ? fmaf(v[u].w, w[u], acc.w) : acc.w;
                } else {
                    acc.x = keep ? acc.x + v[u].x : acc.x;
                    acc.y = keep ? acc.y + v[u].y : acc.y;
                    acc.z = keep ? acc.z + v[u].z : acc.z;
                    acc.w = keep ? acc.w + v[u].w : acc.w;
                }
            }
            if (last) break;
            e += B;
        }
        if (!P::kWeighted) { acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale; }
        *reinterpret_cast<float4 *>(out + (size_t)j * c_out + ch * 4) = acc;
        if (jn >= n_rows) break;
        j = jn; e = ne; e_end = ne_end;
        jn = jnn; ne = nne; ne_end = nne_end;
    }
}

// Two adjacent channel chunks per thread (32 bytes of the gathered row, ONE weight): for policies whose weight
// is shared by an even number of chunks (GVA: I = C/G = 8 channels per group -> the thread owns a whole group).
// Per gathered 16 bytes this halves the perm / weight / address instructions of csr_walk_kernel — the r01h study
// found that walk bound by issue slots (48 %) and the L1 pipe (67 %), not by L2 or DRAM.  Same entry order and
// the same fmaf sequence per channel as csr_walk_kernel: bitwise-identical results.
// `pairs` = chunks / 2 (threads per row); the grid is a multiple of pairs / gcd(pairs, kWalkBlock) CTAs.
template <int B, class P>
__global__ void __launch_bounds__(kWalkBlock)
csr_walk2_kernel(long long n_rows, int pairs, int c_out, const int *__restrict__ rowptr,
                 const int *__restrict__ perm, P pol, float *__restrict__ out) {
    const long long step_items = (long long)gridDim.x * kWalkBlock;  // multiple of pairs (walk_grid)
    const long long t0 = (long long)blockIdx.x * kWalkBlock + threadIdx.x;
    const long long row_step = step_items / pairs;
    long long j = t0 / pairs;
    const int ch = 2 * (int)(t0 - j * pairs);  // first of the thread's two chunks
    if (j >= n_rows) return;

    int e = __ldg(rowptr + j), e_end = __ldg(rowptr + j + 1);
    int pn[B];
#pragma unroll
    for (int u = 0; u < B; ++u) pn[u] = (e + u < e_end) ? __ldg(perm + e + u) : 0;
    long long jn = j + row_step;
    int ne = 0, ne_end = 0;
    if (jn < n_rows) { ne = __ldg(rowptr + jn); ne_end = __ldg(rowptr + jn + 1); }

    for (;;) {
        const long long jnn = jn + row_step;
        int nne = 0, nne_end = 0;
        if (jnn < n_rows) { nne = __ldg(rowptr + jnn); nne_end = __ldg(rowptr + jnn + 1); }
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        for (;;) {
            const bool last = e + B >= e_end;
            int p[B];
            float4 v0[B], v1[B];
            float w[B];
#pragma unroll
            for (int u = 0; u < B; ++u) {
                p[u] = pn[u];
                v0[u] = pol.load(p[u], ch);  // unconditional: slots past the row end hold p = 0 (a valid entry)
                v1[u] = pol.load(p[u], ch + 1);
            }
#pragma unroll
            for (int u = 0; u < B; ++u) w[u] = pol.weight(p[u], ch);
            const int pe = last ? ne : e + B, pe_end = last ? ne_end : e_end;
#pragma unroll
            for (int u = 0; u < B; ++u) pn[u] = (pe + u < pe_end) ? __ldg(perm + pe + u) : 0;
            issue_fence();
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const bool keep = e + u < e_end;
                a0.x = keep ? fmaf(v0[u].x, w[u], a0.x) : a0.x;
                a0.y = keep ? fmaf(v0[u].y, w[u], a0.y) : a0.y;
                a0.z = keep ? fmaf(v0[u].z, w[u], a0.z) : a0.z;
                a0.w = keep ? fmaf(v0[u].w, w[u], a0.w) : a0.w;
                a1.x = keep ? fmaf(v1[u].x, w[u], a1.x) : a1.x;
                a1.y = keep ? fmaf(v1[u].y, w[u], a1.y) : a1.y;
                a1.z = keep ? fmaf(v1[u].z, w[u], a1.z) : a1.z;
                a1.w = keep ? fmaf(v1[u].w, w[u], a1.w) : a1.w;
            }
            if (last) break;
            e += B;
        }
        float *o = out + (size_t)j * c_out + ch * 4;
        *reinterpret_cast<float4 *>(o) = a0;
        *reinterpret_cast<float4 *>(o + 4) = a1;
        if (jn >= n_rows) break;
        j = jn; e = ne; e_end = ne_end;
        jn = jnn; ne = nne; ne_end = nne_end;
    }
}

}  // namespace aopt
