// gva.cu — GroupedVectorAttention tail: softmax over the k neighbours, mask, and the grouped
// weighted sum, fused with the neighbour gather of `value`.
//
// Replaces the torch op chain of
// /root/reference/pointcept/models/point_transformer_v2/point_transformer_v2m2_base.py:
//   :110  value = grouping(idx, value)            (N,k,C) materialised
//   :119  value = value + peb                     (N,k,C) again
//   :122  softmax(weight, dim=1)                  (N,k,G)
//   :124-125 mask = sign(idx+1); weight *= mask
//   :126-128 einsum("n s g i, n s g -> n g i")
// (~8 passes over (N,k,C)/(N,k,G) tensors, each kept for autograd) with one forward and two backward
// kernels that read peb once, gather value rows straight from the (N,C) tensor and never
// materialise the gathered value.  Group layout is PTv2's: channel ch belongs to group ch / (C/G)
// (the reference's own CUDA `aggregation` kernel uses the PTv1 layout ch % w_c and is kept
// separately in legacy.cu).
//
// Thread mapping (128-bit path, I = C/G a multiple of 4): one thread per (point n, 4-channel chunk).
// Consecutive lanes hold consecutive 16-byte chunks of the same (n, s) row, so every warp-level
// load/store of peb / grad_peb covers contiguous memory; the I/4 lanes of one group sit next to each
// other (dot products over a group are finished with 1-2 shuffles).  The neighbour loop is unrolled by
// four with unconditional loads (masked slots read row 0 and are discarded by a select), which keeps
// 8+ independent 128-bit requests in flight per thread — these kernels are latency-bound, not
// issue-bound.  peb / grad_peb stream with L1::no_allocate; value / grad_out rows are re-used by
// neighbouring points and go through the read-only path.  Other widths take the scalar
// one-thread-per-(point, group) kernels (GroupVec<0>).
//
// Algorithmic bytes: forward 4NC + 4NkC + 4NkG(+4NkG prob) + 4Nk + 4NC;
// backward reads 8NC + 4NkC + 4NkG + 8Nk + 4(N+1), writes 4NkC + 4NkG + 4NC.
#include <math.h>
#include <stdlib.h>
#include <initializer_list>

#include "common.cuh"
#include "csr_walk.cuh"

namespace aopt {

constexpr int kGvaBlock = 256;
constexpr int kMaxScalarI = 64;  // scalar fallback: channels of one group held in registers

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_add(const float4 &a, const float4 &b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float dot4(const float4 &a, const float4 &b) {
    return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}
// acc += x * w when keep, else unchanged (select, no branch: masked slots never touch acc)
__device__ __forceinline__ void fma_keep(float4 &acc, const float4 &x, float w, bool keep) {
    acc.x = keep ? fmaf(x.x, w, acc.x) : acc.x;
    acc.y = keep ? fmaf(x.y, w, acc.y) : acc.y;
    acc.z = keep ? fmaf(x.z, w, acc.z) : acc.z;
    acc.w = keep ? fmaf(x.w, w, acc.w) : acc.w;
}

// Sum over the GL adjacent lanes that share one group (GL = I/4 in {1,2,4}); all 32 lanes take part.
template <int GL>
__device__ __forceinline__ float group_sum(float v) {
    if (GL >= 2) v += __shfl_xor_sync(0xffffffffu, v, 1);
    if (GL >= 4) v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

// ---- forward: one thread per (point, 4-channel chunk) ------------------------------------------------
template <int GL>
__global__ void __launch_bounds__(kGvaBlock)
gva_forward_kernel(long long n, int k, int c, int g, const float *__restrict__ value,
                   const float *__restrict__ peb, const float *__restrict__ logits,
                   const int *__restrict__ idx, float *__restrict__ out, float *__restrict__ prob) {
    const int chunks = c >> 2;
    const long long total = n * chunks;
    const long long step = (long long)gridDim.x * kGvaBlock;
    for (long long t = (long long)blockIdx.x * kGvaBlock + threadIdx.x; t < total; t += step) {
        const long long pt = t / chunks;
        const int ch = (int)(t - pt * chunks);
        const int gi = ch / GL;
        const bool writer = (ch % GL) == 0;  // one lane per group stores the probabilities
        const float *lg = logits + (size_t)pt * k * g + gi;
        // softmax over the k neighbour slots of (point, group): exp(x - max) / sum, torch.softmax(dim=1)
        float mx = -INFINITY;
        for (int s = 0; s < k; ++s) mx = fmaxf(mx, __ldg(lg + (size_t)s * g));
        float sum = 0.f;
        for (int s = 0; s < k; ++s) sum += expf(__ldg(lg + (size_t)s * g) - mx);
        const int *ix = idx + (size_t)pt * k;
        const float *vbase = value + ch * 4;
        const float *pe = peb ? peb + (size_t)pt * k * c + ch * 4 : nullptr;
        float *pr = prob ? prob + (size_t)pt * k * g + gi : nullptr;
        float4 acc = f4_zero();
        int s = 0;
        for (; s + 4 <= k; s += 4) {
            int j[4];
            float4 v[4], q[4];
            float p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) j[u] = __ldg(ix + s + u);
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ldg_gather4(vbase + (size_t)max(j[u], 0) * c);
#pragma unroll
            for (int u = 0; u < 4; ++u) q[u] = pe ? ldg_stream4(pe + (size_t)(s + u) * c) : f4_zero();
#pragma unroll
            for (int u = 0; u < 4; ++u) p[u] = expf(__ldg(lg + (size_t)(s + u) * g) - mx) / sum;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (pr && writer) pr[(size_t)(s + u) * g] = p[u];
                fma_keep(acc, f4_add(v[u], q[u]), p[u], j[u] >= 0);  // sign(idx+1) mask
            }
        }
        for (; s < k; ++s) {
            const int j = __ldg(ix + s);
            const float4 v = ldg_gather4(vbase + (size_t)max(j, 0) * c);
            const float4 q = pe ? ldg_stream4(pe + (size_t)s * c) : f4_zero();
            const float p = expf(__ldg(lg + (size_t)s * g) - mx) / sum;
            if (pr && writer) pr[(size_t)s * g] = p;
            fma_keep(acc, f4_add(v, q), p, j >= 0);
        }
        *reinterpret_cast<float4 *>(out + (size_t)pt * c + ch * 4) = acc;
    }
}

// ---- forward, compile-time neighbour count (NS in {8,16,32}) -----------------------------------------
// Same mapping as gva_forward_kernel, restructured around the DRAM latency (the generic kernel exposes it
// ~9 times per point and ptxas sinks register loads next to their uses, leaving ~3 requests in flight):
//   * the idx row of the NEXT item is prefetched into registers while the current one is processed;
//   * at the top of an item all NS peb pieces and all NS gathered value pieces are requested at once
//     with cp.async into the thread's own shared-memory slots (no registers held, no barrier);
//   * the logits column is loaded and the softmax computed while those copies are in flight.
// One latency window per item instead of nine; 2·NS·16 bytes of shared memory per thread.
constexpr int kGvaNsBlock = 128;
// threads per CTA of the NS kernels: 2·NS 16-byte slots per thread — 128 threads at NS <= 16 (64 KB at NS = 16: three
// CTAs per SM); 64 threads at NS = 32, where 128 would need 128 KB and leave ONE CTA (4 warps) per SM: the k = 32
// schedule ran these kernels at 49-57 % of the HBM peak (profiles/r02a, scannet150k k32 variant).
template <int NS> __host__ __device__ constexpr int ns_block() { return NS >= 32 ? 64 : kGvaNsBlock; }

// Gathered value rows.  Measured at level 0 on one box (A/B in the same process): forward 232 us with L2-only
// copies (.cg) vs 244 us with L1-allocating ones (.ca); backward_query 399 us (.cg) vs 368 us (.ca).  Defaults
// follow that; AOPT_GVA_GATHER=<fwd><bwd>[<fused bwd>] with letters g / a overrides (e.g. "gg", "aa", "gaa").
// The FUSED backward takes .cg: its L1 (~60 KB beside the staging slots) is what the CSR walk lives on — sector re-use of
// the perm / probability / idx loads and the grad_out rows of neighbouring queries — and the value rows only evict that:
// 457 -> 440 us at level 0 (profiles/r04d_kernel_bench_l1.txt).  The same run showed how much the walk depends on L1:
// CTA shapes that push the shared-memory carve-out from 196 to 228 KB (L1 60 -> 28 KB) cost 27-76 %, and
// L1::no_allocate on the "streamed" scalar loads 39 % (a 128-byte line serves 5 consecutive slots of a probability column).
__constant__ int g_gva_gather_ca = 2;   // bit 0: forward uses .ca, bit 1: backward_query uses .ca, bit 2: fused backward uses .ca
template <int WHICH>
__device__ __forceinline__ void cp_async16_row(float4 *dst, const float *src) {
    if (g_gva_gather_ca & WHICH) cp_async16_gather(dst, src);
    else cp_async16_stream(dst, src);
}

__device__ __forceinline__ void load_idx_row16(const int *__restrict__ row, int *j, int count4) {
    const int4 *r4 = reinterpret_cast<const int4 *>(row);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        if (u < count4) {
            const int4 q4 = __ldg(r4 + u);
            j[4 * u] = q4.x; j[4 * u + 1] = q4.y; j[4 * u + 2] = q4.z; j[4 * u + 3] = q4.w;
        }
    }
}

template <int GL, int NS, bool HAS_PEB>
__global__ void __launch_bounds__(ns_block<NS>())
gva_forward_ns_kernel(long long n, int c, int g, const float *__restrict__ value,
                      const float *__restrict__ peb, const float *__restrict__ logits,
                      const int *__restrict__ idx, float *__restrict__ out, float *__restrict__ prob, int pf) {
    constexpr int BLK = ns_block<NS>();
    extern __shared__ float4 stage[];  // [2*NS][BLK]: value pieces, then peb pieces
    float4 *sv = stage + threadIdx.x;
    float4 *sq = stage + NS * BLK + threadIdx.x;
    const int chunks = c >> 2;
    const long long total = n * chunks;
    const long long step = (long long)gridDim.x * BLK;
    long long t = (long long)blockIdx.x * BLK + threadIdx.x;
    int jn[NS];
    if (t < total) load_idx_row16(idx + (size_t)(t / chunks) * NS, jn, NS / 4);
    for (; t < total; t += step) {
        const long long pt = t / chunks;
        const int ch = (int)(t - pt * chunks);
        const int gi = ch / GL;
        const bool writer = (ch % GL) == 0;
        int j[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) j[s] = jn[s];
        const float *vbase = value + ch * 4;
#pragma unroll
        for (int s = 0; s < NS; ++s) cp_async16_row<1>(sv + s * BLK, vbase + (size_t)max(j[s], 0) * c);
        if (HAS_PEB) {
            const float *pe = peb + (size_t)pt * NS * c + ch * 4;
#pragma unroll
            for (int s = 0; s < NS; ++s) cp_async16_stream(sq + s * BLK, pe + (size_t)s * c);
        }
        cp_async_commit();
        const float *lg = logits + (size_t)pt * NS * g + gi;
        float e[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) e[s] = __ldg(lg + (size_t)s * g);
        if (t + step < total) {
            const long long ptn = (t + step) / chunks;
            load_idx_row16(idx + (size_t)ptn * NS, jn, NS / 4);
            // tuning "l2pf": the NEXT item's peb block (NS·c floats, contiguous per point) is requested into L2 now, by the lane
            // that owns the point's first chunk; its cp.async copies one iteration later are L2 hits.  Neutral here (the
            // forward is at 93 % of the HBM peak either way), kept for symmetry with the fused backward, where it pays.
            if (HAS_PEB && pf && t + step == ptn * chunks) l2_prefetch_bulk(peb + (size_t)ptn * NS * c, (unsigned)(NS * c * 4));
        }
        float mx = -INFINITY;
#pragma unroll
        for (int s = 0; s < NS; ++s) mx = fmaxf(mx, e[s]);
        float sum = 0.f;
#pragma unroll
        for (int s = 0; s < NS; ++s) { e[s] = expf(e[s] - mx); sum += e[s]; }
        // softmax probabilities: one IEEE division, then multiplies (within 1 ulp of exp(x-max)/sum)
        const float inv = 1.0f / sum;
#pragma unroll
        for (int s = 0; s < NS; ++s) e[s] *= inv;
        if (prob && writer) {
            float *pr = prob + (size_t)pt * NS * g + gi;
#pragma unroll
            for (int s = 0; s < NS; ++s) pr[(size_t)s * g] = e[s];
        }
        cp_async_wait_all();
        float4 acc = f4_zero();
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const float4 v = sv[s * BLK];
            const float4 q = HAS_PEB ? sq[s * BLK] : f4_zero();
            fma_keep(acc, f4_add(v, q), e[s], j[s] >= 0);  // sign(idx+1) mask
        }
        *reinterpret_cast<float4 *>(out + (size_t)pt * c + ch * 4) = acc;
    }
}

// ---- backward, per query: grad_peb and grad_logits ---------------------------------------------------
// grad_logits[n,s,g] = p_s (gw_s - Σ_s' p_s' gw_s'),  gw_s = mask_s <grad_out[n, group g], value[idx]+peb>.
template <int GL>
__global__ void __launch_bounds__(kGvaBlock)
gva_backward_query_kernel(long long n, int k, int c, int g, const float *__restrict__ grad_out,
                          const float *__restrict__ value, const float *__restrict__ peb,
                          const float *__restrict__ prob, const int *__restrict__ idx,
                          float *__restrict__ grad_peb, float *grad_logits) {
    const int chunks = c >> 2;
    const long long total = n * chunks;  // a multiple of GL, so the lanes of a group are in or out together
    const long long step = (long long)gridDim.x * kGvaBlock;
    // warp-uniform trip count: lanes past the end run on a clamped index and store nothing, so the
    // group shuffles below always see all 32 lanes
    for (long long base = (long long)blockIdx.x * kGvaBlock + (threadIdx.x & ~31); base < total; base += step) {
        const long long t_raw = base + (threadIdx.x & 31);
        const bool live = t_raw < total;
        const long long t = live ? t_raw : total - 1;
        const long long pt = t / chunks;
        const int ch = (int)(t - pt * chunks);
        const int gi = ch / GL;
        const bool writer = live && (ch % GL) == 0;
        const float4 go = ldg_gather4(grad_out + (size_t)pt * c + ch * 4);
        const int *ix = idx + (size_t)pt * k;
        const float *vbase = value + ch * 4;
        const float *pe = peb ? peb + (size_t)pt * k * c + ch * 4 : nullptr;
        float *gp = (grad_peb && live) ? grad_peb + (size_t)pt * k * c + ch * 4 : nullptr;
        const float *pr = prob + (size_t)pt * k * g + gi;
        float *gl = grad_logits + (size_t)pt * k * g + gi;
        float dot = 0.f;  // Σ_s p_s · dL/dp_s
        int s = 0;
        for (; s + 4 <= k; s += 4) {
            int j[4];
            float4 v[4], q[4];
            float p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) j[u] = __ldg(ix + s + u);
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ldg_gather4(vbase + (size_t)max(j[u], 0) * c);
#pragma unroll
            for (int u = 0; u < 4; ++u) q[u] = pe ? ldg_stream4(pe + (size_t)(s + u) * c) : f4_zero();
#pragma unroll
            for (int u = 0; u < 4; ++u) p[u] = __ldg(pr + (size_t)(s + u) * g);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool keep = j[u] >= 0;
                float gw = group_sum<GL>(dot4(go, f4_add(v[u], q[u])));
                gw = keep ? gw : 0.f;
                const float w = keep ? p[u] : 0.f;
                if (gp) stg_stream4(gp + (size_t)(s + u) * c, make_float4(go.x * w, go.y * w, go.z * w, go.w * w));
                dot = fmaf(p[u], gw, dot);
                if (writer) gl[(size_t)(s + u) * g] = gw;  // parked; finalised below by the same thread
            }
        }
        for (; s < k; ++s) {
            const int j = __ldg(ix + s);
            const bool keep = j >= 0;
            const float4 v = ldg_gather4(vbase + (size_t)max(j, 0) * c);
            const float4 q = pe ? ldg_stream4(pe + (size_t)s * c) : f4_zero();
            const float p = __ldg(pr + (size_t)s * g);
            float gw = group_sum<GL>(dot4(go, f4_add(v, q)));
            gw = keep ? gw : 0.f;
            const float w = keep ? p : 0.f;
            if (gp) stg_stream4(gp + (size_t)s * c, make_float4(go.x * w, go.y * w, go.z * w, go.w * w));
            dot = fmaf(p, gw, dot);
            if (writer) gl[(size_t)s * g] = gw;
        }
        if (writer) {
            for (s = 0; s < k; ++s) {
                const float p = __ldg(pr + (size_t)s * g);
                gl[(size_t)s * g] = p * (gl[(size_t)s * g] - dot);  // softmax backward
            }
        }
    }
}

// Compile-time neighbour count: same latency-oriented structure as gva_forward_ns_kernel (idx prefetch,
// all peb / value pieces requested at once with cp.async into per-thread slots); the probability
// column is loaded while the copies fly and gw stays in registers (no parking in grad_logits).
template <int GL, int NS, bool HAS_PEB>
__global__ void __launch_bounds__(ns_block<NS>())
gva_backward_query_ns_kernel(long long n, int c, int g, const float *__restrict__ grad_out,
                             const float *__restrict__ value, const float *__restrict__ peb,
                             const float *__restrict__ prob, const int *__restrict__ idx,
                             float *__restrict__ grad_peb, float *__restrict__ grad_logits) {
    constexpr int BLK = ns_block<NS>();
    extern __shared__ float4 stage[];
    float4 *sv = stage + threadIdx.x;
    float4 *sq = stage + NS * BLK + threadIdx.x;
    const int chunks = c >> 2;
    const long long total = n * chunks;
    const long long step = (long long)gridDim.x * BLK;
    // warp-uniform trip count (see gva_backward_query_kernel)
    long long base = (long long)blockIdx.x * BLK + (threadIdx.x & ~31);
    const int lane = threadIdx.x & 31;
    int jn[NS];
    if (base < total) load_idx_row16(idx + (size_t)(min(base + lane, total - 1) / chunks) * NS, jn, NS / 4);
    for (; base < total; base += step) {
        const long long t_raw = base + lane;
        const bool live = t_raw < total;
        const long long t = live ? t_raw : total - 1;
        const long long pt = t / chunks;
        const int ch = (int)(t - pt * chunks);
        const int gi = ch / GL;
        const bool writer = live && (ch % GL) == 0;
        int j[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) j[s] = jn[s];
        const float *vbase = value + ch * 4;
#pragma unroll
        for (int s = 0; s < NS; ++s) cp_async16_row<2>(sv + s * BLK, vbase + (size_t)max(j[s], 0) * c);
        if (HAS_PEB) {
            const float *pe = peb + (size_t)pt * NS * c + ch * 4;
#pragma unroll
            for (int s = 0; s < NS; ++s) cp_async16_stream(sq + s * BLK, pe + (size_t)s * c);
        }
        cp_async_commit();
        const float4 go = ldg_gather4(grad_out + (size_t)pt * c + ch * 4);
        const float *pr = prob + (size_t)pt * NS * g + gi;
        float p[NS], gw[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) p[s] = __ldg(pr + (size_t)s * g);
        if (base + step < total)
            load_idx_row16(idx + (size_t)(min(base + step + lane, total - 1) / chunks) * NS, jn, NS / 4);
        float *gp = (HAS_PEB && grad_peb && live) ? grad_peb + (size_t)pt * NS * c + ch * 4 : nullptr;
        cp_async_wait_all();
        float dot = 0.f;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const bool keep = j[s] >= 0;
            const float4 v = sv[s * BLK];
            const float4 q = HAS_PEB ? sq[s * BLK] : f4_zero();
            const float d = group_sum<GL>(dot4(go, f4_add(v, q)));
            gw[s] = keep ? d : 0.f;
            const float w = keep ? p[s] : 0.f;
            if (gp) stg_stream4(gp + (size_t)s * c, make_float4(go.x * w, go.y * w, go.z * w, go.w * w));
            dot = fmaf(p[s], gw[s], dot);
        }
        if (writer) {
            float *gl = grad_logits + (size_t)pt * NS * g + gi;
#pragma unroll
            for (int s = 0; s < NS; ++s) gl[(size_t)s * g] = p[s] * (gw[s] - dot);  // softmax backward
        }
    }
}

// ---- backward, fused: grad_peb + grad_logits (per query) AND grad_value (per source) in one kernel -------
// In self-attention the queries ARE the sources (n_src == n), so the thread that owns (point, chunk) in the
// query pass above also owns the same (source row, chunk) of the CSR walk of gva_backward_value.  The two
// halves bound each other's idle resource: the query half waits on DRAM (HBM-bound: 32 streamed / gathered
// 16-byte pieces per item, cp.async), the walk half is a dependent chain of L2 hits (rowptr -> perm ->
// grad_out row + probability, latency-bound at ~30 % of the HBM pipe when it runs alone).  Here the walk of a
// thread's row runs BETWEEN issuing the item's cp.async copies and waiting for them, i.e. in the shadow of
// the DRAM latency of the same thread.  DRAM traffic = SURVEY §8d's "fused GVA bwd": the second read of prob /
// grad_out (by the walk) is an L2 hit because the grid sweeps one compact window of rows.
// Same per-row entry order (ascending p) and the same fmaf sequence as csr_walk_kernel<8, BvPolicy>: grad_value
// is bitwise identical to the two-kernel path; grad_peb / grad_logits are the code of the kernel above.
constexpr int kFusedBatch = 8;

template <int GL, int NS, bool HAS_PEB>
__global__ void __launch_bounds__(ns_block<NS>())
gva_backward_fused_ns_kernel(long long n, int c, int g, const float *__restrict__ grad_out,
                             const float *__restrict__ value, const float *__restrict__ peb,
                             const float *__restrict__ prob, const int *__restrict__ idx,
                             const int *__restrict__ rowptr, const int *__restrict__ perm,
                             float *__restrict__ grad_peb, float *__restrict__ grad_logits,
                             float *__restrict__ grad_value, int pf) {
    constexpr int BLK = ns_block<NS>();
    extern __shared__ float4 stage[];
    constexpr int B = kFusedBatch;
    constexpr int KSHIFT = NS == 8 ? 3 : NS == 16 ? 4 : 5;
    float4 *sv = stage + threadIdx.x;
    float4 *sq = stage + NS * BLK + threadIdx.x;
    const int chunks = c >> 2;
    const long long total = n * chunks;
    const long long step = (long long)gridDim.x * BLK;
    long long base = (long long)blockIdx.x * BLK + (threadIdx.x & ~31);
    const int lane = threadIdx.x & 31;
    int jn[NS];
    // walk state: [e, e_end) = CSR row of the current item, pn = its first B perm values, (ne, ne_end) = next item's row
    int e = 0, e_end = 0, ne = 0, ne_end = 0;
    int pn[B];
    if (base < total) {
        const long long pt0 = min(base + lane, total - 1) / chunks;
        load_idx_row16(idx + (size_t)pt0 * NS, jn, NS / 4);
        e = __ldg(rowptr + pt0); e_end = __ldg(rowptr + pt0 + 1);
#pragma unroll
        for (int u = 0; u < B; ++u) pn[u] = (e + u < e_end) ? __ldg(perm + e + u) : 0;
        if (base + step < total) {
            const long long pt1 = min(base + step + lane, total - 1) / chunks;
            ne = __ldg(rowptr + pt1); ne_end = __ldg(rowptr + pt1 + 1);
        }
    }
    for (; base < total; base += step) {
        const long long t_raw = base + lane;
        const bool live = t_raw < total;
        const long long t = live ? t_raw : total - 1;
        const long long pt = t / chunks;
        const int ch = (int)(t - pt * chunks);
        const int gi = ch / GL;
        const bool writer = live && (ch % GL) == 0;
        int j[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) j[s] = jn[s];
        const float *vbase = value + ch * 4;
#pragma unroll
        for (int s = 0; s < NS; ++s) cp_async16_row<4>(sv + s * BLK, vbase + (size_t)max(j[s], 0) * c);
        if (HAS_PEB) {
            const float *pe = peb + (size_t)pt * NS * c + ch * 4;
#pragma unroll
            for (int s = 0; s < NS; ++s) cp_async16_stream(sq + s * BLK, pe + (size_t)s * c);
        }
        cp_async_commit();
        const float4 go = ldg_gather4(grad_out + (size_t)pt * c + ch * 4);
        const float *pr = prob + (size_t)pt * NS * g + gi;
        float p[NS], gw[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) p[s] = __ldg(pr + (size_t)s * g);
        const bool more = base + step < total;
        if (more) {
            const long long tn = min(base + step + lane, total - 1);
            const long long ptn = tn / chunks;
            load_idx_row16(idx + (size_t)ptn * NS, jn, NS / 4);
            // tuning "l2pf" (default on): the NEXT item's peb block goes from DRAM to L2 while this item's walk runs.  A warp
            // spends most of an item in the walk's dependent L2 round trips with nothing of its own in flight towards DRAM
            // (12 warps per SM: ~27 KB in flight per SM on average, below what the HBM pipe needs); the bulk prefetch (copy
            // engine, no LSU slot, no register, no shared memory) keeps the DRAM stream going: 480 -> 461 us at level 0,
            // 149 -> 139 us at level 1, 80 -> 74 us at level 2 (profiles/r04b_kernel_bench_l2pf.txt).
            if (HAS_PEB && pf && tn == ptn * chunks && base + step + lane < total)
                l2_prefetch_bulk(peb + (size_t)ptn * NS * c, (unsigned)(NS * c * 4));
        }
        // rowptr of the item after next (two rows ahead, like csr_walk_kernel)
        int nne = 0, nne_end = 0;
        if (base + 2 * step < total) {
            const long long pt2 = min(base + 2 * step + lane, total - 1) / chunks;
            nne = __ldg(rowptr + pt2); nne_end = __ldg(rowptr + pt2 + 1);
        }
        // ---- CSR walk of source row `pt`, chunk `ch` (grad_value), under the copies issued above ----
        {
            float4 acc = f4_zero();
            const float *gbase = grad_out + ch * 4;
            const float *wbase = prob + gi;
            for (;;) {
                const bool last = e + B >= e_end;
                int q[B];
                float4 v[B];
                float w[B];
#pragma unroll
                for (int u = 0; u < B; ++u) {
                    q[u] = pn[u];   // flat (query, slot) position; slots past the row end hold 0 (a valid entry)
                    v[u] = ldg_gather4(gbase + (size_t)(q[u] >> KSHIFT) * c);
                }
#pragma unroll
                for (int u = 0; u < B; ++u) w[u] = __ldg(wbase + (size_t)q[u] * g);
                const int pe0 = last ? ne : e + B, pe1 = last ? ne_end : e_end;
#pragma unroll
                for (int u = 0; u < B; ++u) pn[u] = (pe0 + u < pe1) ? __ldg(perm + pe0 + u) : 0;
                issue_fence();
#pragma unroll
                for (int u = 0; u < B; ++u) fma_keep(acc, v[u], w[u], e + u < e_end);
                if (last) break;
                e += B;
            }
            if (live) *reinterpret_cast<float4 *>(grad_value + (size_t)pt * c + ch * 4) = acc;
            e = ne; e_end = ne_end; ne = nne; ne_end = nne_end;
        }
        // ---- query half (identical to gva_backward_query_ns_kernel) ----
        float *gp = (HAS_PEB && grad_peb && live) ? grad_peb + (size_t)pt * NS * c + ch * 4 : nullptr;
        cp_async_wait_all();
        float dot = 0.f;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const bool keep = j[s] >= 0;
            const float4 v = sv[s * BLK];
            const float4 q = HAS_PEB ? sq[s * BLK] : f4_zero();
            const float d = group_sum<GL>(dot4(go, f4_add(v, q)));
            gw[s] = keep ? d : 0.f;
            const float w = keep ? p[s] : 0.f;
            if (gp) stg_stream4(gp + (size_t)s * c, make_float4(go.x * w, go.y * w, go.z * w, go.w * w));
            dot = fmaf(p[s], gw[s], dot);
        }
        if (writer) {
            float *gl = grad_logits + (size_t)pt * NS * g + gi;
#pragma unroll
            for (int s = 0; s < NS; ++s) gl[(size_t)s * g] = p[s] * (gw[s] - dot);  // softmax backward
        }
    }
}

// ---- backward, per source: grad_value through the CSR ------------------------------------------------
// kshift >= 0: k is a power of two and q = p >> kshift; otherwise q = p / k.
// The row walk is a dependent chain (rowptr → perm → prob / grad_out), so entries are taken eight at a
// time: eight perm loads, then eight grad_out pieces requested with cp.async into the thread's own
// shared-memory slots plus eight probability loads, with the next eight perm values prefetched while
// those are in flight.
constexpr int kBvBlock = 256;
constexpr int kBvBatch = 8;

template <int GL>
__global__ void __launch_bounds__(kBvBlock)
gva_backward_value_kernel(long long n_src, int k, int kshift, int c, int g, const float *__restrict__ grad_out,
                          const float *__restrict__ prob, const int *__restrict__ rowptr,
                          const int *__restrict__ perm, float *__restrict__ grad_value) {
    __shared__ float4 stage[kBvBatch * kBvBlock];
    float4 *sg = stage + threadIdx.x;
    const int chunks = c >> 2;
    const long long total = n_src * chunks;
    const long long step = (long long)gridDim.x * kBvBlock;
    for (long long t = (long long)blockIdx.x * kBvBlock + threadIdx.x; t < total; t += step) {
        const long long j = t / chunks;
        const int ch = (int)(t - j * chunks);
        const int gi = ch / GL;
        const float *gbase = grad_out + ch * 4;
        float4 acc = f4_zero();
        int e = __ldg(rowptr + j);
        const int e_end = __ldg(rowptr + j + 1);
        int pn[kBvBatch];
#pragma unroll
        for (int u = 0; u < kBvBatch; ++u) pn[u] = (e + u < e_end) ? __ldg(perm + e + u) : 0;
        for (; e < e_end; e += kBvBatch) {
            int p[kBvBatch];
            float w[kBvBatch];
#pragma unroll
            for (int u = 0; u < kBvBatch; ++u) {
                p[u] = pn[u];  // flat (query, slot) position; idx[p] == j >= 0, so the mask is 1
                const int q = kshift >= 0 ? (p[u] >> kshift) : (p[u] / k);
                cp_async16_gather(sg + u * kBvBlock, gbase + (size_t)q * c);
            }
            cp_async_commit();
#pragma unroll
            for (int u = 0; u < kBvBatch; ++u) w[u] = (e + u < e_end) ? __ldg(prob + (size_t)p[u] * g + gi) : 0.f;
#pragma unroll
            for (int u = 0; u < kBvBatch; ++u)
                pn[u] = (e + kBvBatch + u < e_end) ? __ldg(perm + e + kBvBatch + u) : 0;
            cp_async_wait_all();
#pragma unroll
            for (int u = 0; u < kBvBatch; ++u) fma_keep(acc, sg[u * kBvBlock], w[u], e + u < e_end);
        }
        *reinterpret_cast<float4 *>(grad_value + (size_t)j * c + ch * 4) = acc;
    }
}

// csr_walk.cuh policy: entry p = flat (query, slot); gathers grad_out[query] (re-used by the k sources of a
// query: read-only path, L1-allocating) and the probability of (query, slot, group of the chunk).
template <int GL>
struct BvPolicy {
    const float *grad_out, *prob;
    int c, g, k, kshift;
    static constexpr bool kWeighted = true;
    __device__ __forceinline__ float4 load(int p, int ch) const {
        const int q = kshift >= 0 ? (p >> kshift) : (p / k);
        return ldg_gather4(grad_out + (size_t)q * c + ch * 4);
    }
    __device__ __forceinline__ float weight(int p, int ch) const { return __ldg(prob + (size_t)p * g + ch / GL); }
};

template <int GL>
static void launch_bv_walk(long long n_src, int k, int c, int g, const float *grad_out, const float *prob,
                           const int *rowptr, const int *perm, float *grad_value, cudaStream_t st) {
    const int chunks = c / 4;
    const BvPolicy<GL> pol{grad_out, prob, c, g, k, 0};
    BvPolicy<GL> p2 = pol;
    int sh = -1;
    for (int b = 0; b < 31; ++b) if ((1 << b) == k) sh = b;
    p2.kshift = sh;
    if (walk_batch() == 4)
        csr_walk_kernel<4, BvPolicy<GL>><<<walk_grid(n_src, chunks, 12), kWalkBlock, 0, st>>>(n_src, chunks, c, rowptr, perm, p2, 1.f, grad_value);
    else
        csr_walk_kernel<8, BvPolicy<GL>><<<walk_grid(n_src, chunks, 12), kWalkBlock, 0, st>>>(n_src, chunks, c, rowptr, perm, p2, 1.f, grad_value);
}

// Group-per-thread walk (csr_walk2_kernel): the thread owns both chunks of an 8-channel group, one probability
// per entry.  AOPT_BV_IMPL=group / chunk selects it or the chunk-per-thread walk (A/B measurements); default below.
static int bv_group_mode() {  // 1 = group walk where it applies (GL == 2), 0 = chunk walk
    static const int m = [] {
        const char *e = getenv("AOPT_BV_IMPL");
        if (e && e[0] == 'g') return 1;
        if (e && e[0] == 'c') return 0;
        return 0;
    }();
    return m;
}
static int bv_group_batch() {  // entries per batch of the group walk: 4 (default) or 8 (AOPT_BV_GROUP_B=8)
    static const int b = [] { const char *e = getenv("AOPT_BV_GROUP_B"); return (e && e[0] == '8') ? 8 : 4; }();
    return b;
}
static void launch_bv_group_walk(long long n_src, int k, int c, int g, const float *grad_out, const float *prob,
                                 const int *rowptr, const int *perm, float *grad_value, cudaStream_t st) {
    const int pairs = c / 8;  // == g when I == 8
    BvPolicy<2> pol{grad_out, prob, c, g, k, -1};
    for (int b = 0; b < 31; ++b) if ((1 << b) == k) pol.kshift = b;
    if (bv_group_batch() == 8)
        csr_walk2_kernel<8, BvPolicy<2>><<<walk_grid(n_src, pairs, 8), kWalkBlock, 0, st>>>(n_src, pairs, c, rowptr, perm, pol, grad_value);
    else
        csr_walk2_kernel<4, BvPolicy<2>><<<walk_grid(n_src, pairs, 12), kWalkBlock, 0, st>>>(n_src, pairs, c, rowptr, perm, pol, grad_value);
}

// Four entries at a time with plain loads (tuning alternative, AOPT_BV_IMPL=unroll4).
template <int GL>
__global__ void __launch_bounds__(kGvaBlock)
gva_backward_value_unroll4_kernel(long long n_src, int k, int kshift, int c, int g, const float *__restrict__ grad_out,
                          const float *__restrict__ prob, const int *__restrict__ rowptr,
                          const int *__restrict__ perm, float *__restrict__ grad_value) {
    const int chunks = c >> 2;
    const long long total = n_src * chunks;
    const long long step = (long long)gridDim.x * kGvaBlock;
    for (long long t = (long long)blockIdx.x * kGvaBlock + threadIdx.x; t < total; t += step) {
        const long long j = t / chunks;
        const int ch = (int)(t - j * chunks);
        const int gi = ch / GL;
        const float *gbase = grad_out + ch * 4;
        float4 acc = f4_zero();
        int e = __ldg(rowptr + j);
        const int e_end = __ldg(rowptr + j + 1);
        for (; e + 4 <= e_end; e += 4) {
            int p[4];
            float w[4];
            float4 go[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) p[u] = __ldg(perm + e + u);  // flat (query, slot); idx[p] == j >= 0, mask = 1
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int q = kshift >= 0 ? (p[u] >> kshift) : (p[u] / k);
                go[u] = ldg_gather4(gbase + (size_t)q * c);
                w[u] = __ldg(prob + (size_t)p[u] * g + gi);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) fma_keep(acc, go[u], w[u], true);
        }
        for (; e < e_end; ++e) {
            const int p = __ldg(perm + e);
            const int q = kshift >= 0 ? (p >> kshift) : (p / k);
            fma_keep(acc, ldg_gather4(gbase + (size_t)q * c), __ldg(prob + (size_t)p * g + gi), true);
        }
        *reinterpret_cast<float4 *>(grad_value + (size_t)j * c + ch * 4) = acc;
    }
}

// ---- scalar fallback (I not a multiple of 4, I not in {4,8,16}, or unaligned pointers) ----------------
// One thread per (point, group) holding the I channels of the group in registers.
__global__ void __launch_bounds__(kGvaBlock)
gva_forward_scalar_kernel(long long n, int k, int c, int g, int I, const float *__restrict__ value,
                          const float *__restrict__ peb, const float *__restrict__ logits,
                          const int *__restrict__ idx, float *__restrict__ out, float *__restrict__ prob) {
    const long long total = n * g;
    const long long step = (long long)gridDim.x * kGvaBlock;
    for (long long t = (long long)blockIdx.x * kGvaBlock + threadIdx.x; t < total; t += step) {
        const long long pt = t / g;
        const int gi = (int)(t - pt * g);
        const float *lg = logits + (size_t)pt * k * g + gi;
        float mx = -INFINITY;
        for (int s = 0; s < k; ++s) mx = fmaxf(mx, __ldg(lg + (size_t)s * g));
        float sum = 0.f;
        for (int s = 0; s < k; ++s) sum += expf(__ldg(lg + (size_t)s * g) - mx);
        float acc[kMaxScalarI];
        for (int i = 0; i < I; ++i) acc[i] = 0.f;
        const size_t ch0 = (size_t)gi * I;
        for (int s = 0; s < k; ++s) {
            const float p = expf(__ldg(lg + (size_t)s * g) - mx) / sum;
            if (prob) prob[((size_t)pt * k + s) * g + gi] = p;
            const int j = __ldg(idx + (size_t)pt * k + s);
            if (j < 0) continue;
            for (int i = 0; i < I; ++i) {
                float v = __ldg(value + (size_t)j * c + ch0 + i);
                if (peb) v += __ldg(peb + ((size_t)pt * k + s) * c + ch0 + i);
                acc[i] = fmaf(v, p, acc[i]);
            }
        }
        for (int i = 0; i < I; ++i) out[(size_t)pt * c + ch0 + i] = acc[i];
    }
}

__global__ void __launch_bounds__(kGvaBlock)
gva_backward_query_scalar_kernel(long long n, int k, int c, int g, int I, const float *__restrict__ grad_out,
                                 const float *__restrict__ value, const float *__restrict__ peb,
                                 const float *__restrict__ prob, const int *__restrict__ idx,
                                 float *__restrict__ grad_peb, float *grad_logits) {
    const long long total = n * g;
    const long long step = (long long)gridDim.x * kGvaBlock;
    for (long long t = (long long)blockIdx.x * kGvaBlock + threadIdx.x; t < total; t += step) {
        const long long pt = t / g;
        const int gi = (int)(t - pt * g);
        const size_t ch0 = (size_t)gi * I;
        const float *go = grad_out + (size_t)pt * c + ch0;
        const float *pr = prob + (size_t)pt * k * g + gi;
        float *gl = grad_logits + (size_t)pt * k * g + gi;
        float dot = 0.f;
        for (int s = 0; s < k; ++s) {
            const float p = __ldg(pr + (size_t)s * g);
            const int j = __ldg(idx + (size_t)pt * k + s);
            float gw = 0.f;
            for (int i = 0; i < I; ++i) {
                const float gv = __ldg(go + i);
                if (j >= 0) {
                    float v = __ldg(value + (size_t)j * c + ch0 + i);
                    if (peb) v += __ldg(peb + ((size_t)pt * k + s) * c + ch0 + i);
                    gw = fmaf(gv, v, gw);
                }
                if (grad_peb) grad_peb[((size_t)pt * k + s) * c + ch0 + i] = gv * (j >= 0 ? p : 0.f);
            }
            dot = fmaf(p, gw, dot);
            gl[(size_t)s * g] = gw;
        }
        for (int s = 0; s < k; ++s) {
            const float p = __ldg(pr + (size_t)s * g);
            gl[(size_t)s * g] = p * (gl[(size_t)s * g] - dot);
        }
    }
}

__global__ void __launch_bounds__(kGvaBlock)
gva_backward_value_scalar_kernel(long long n_src, int k, int c, int g, int I, const float *__restrict__ grad_out,
                                 const float *__restrict__ prob, const int *__restrict__ rowptr,
                                 const int *__restrict__ perm, float *__restrict__ grad_value) {
    const long long total = n_src * g;
    const long long step = (long long)gridDim.x * kGvaBlock;
    for (long long t = (long long)blockIdx.x * kGvaBlock + threadIdx.x; t < total; t += step) {
        const long long j = t / g;
        const int gi = (int)(t - j * g);
        const size_t ch0 = (size_t)gi * I;
        float acc[kMaxScalarI];
        for (int i = 0; i < I; ++i) acc[i] = 0.f;
        const int e_end = __ldg(rowptr + j + 1);
        for (int e = __ldg(rowptr + j); e < e_end; ++e) {
            const int p = __ldg(perm + e);
            const int q = p / k;
            const float w = __ldg(prob + (size_t)p * g + gi);
            for (int i = 0; i < I; ++i) acc[i] = fmaf(__ldg(grad_out + (size_t)q * c + ch0 + i), w, acc[i]);
        }
        for (int i = 0; i < I; ++i) grad_value[(size_t)j * c + ch0 + i] = acc[i];
    }
}

// Lanes per group for the 128-bit path: I in {4,8,16} and 16-byte aligned pointers; 0 = scalar path.
static int pick_gl(int c, int I, std::initializer_list<const void *> ptrs) {
    if (I != 4 && I != 8 && I != 16) return 0;
    if (c % 4 != 0) return 0;
    for (const void *p : ptrs)
        if (p && !aligned16(p)) return 0;
    return I / 4;
}

static int log2_exact(int k) {
    for (int sft = 0; sft < 31; ++sft)
        if ((1 << sft) == k) return sft;
    return -1;
}

}  // namespace aopt

using namespace aopt;

// Grid of the NS kernels: GRID here is the work-item count; resident CTAs per SM follow from the
// 2·NS·16·128 bytes of shared memory each CTA needs (227 KB per SM).
static void gva_gather_mode_init() {
    static const bool once = [] {
        const char *e = getenv("AOPT_GVA_GATHER");
        if (e && e[0] && e[1]) {
            const int ca = (e[0] == 'a' ? 1 : 0) | (e[1] == 'a' ? 2 : 0) | (e[2] == 'a' ? 4 : 0);
            cudaMemcpyToSymbol(g_gva_gather_ca, &ca, sizeof(int));
        }
        return true;
    }();
    (void)once;
}

static int ns_grid(long long items, int ns, int block) {
    gva_gather_mode_init(); return stride_grid(items, block, ns <= 8 ? 6 : 3); }

// Dynamic shared memory of the NS kernels: 2·NS slots of 16 bytes per thread (64 KB at NS=16 → opt-in).
#define GVA_DISPATCH_NS2(GLV, NSV, PEB, KERNEL, GRID, ST, ...)                                          \
    {                                                                                                  \
        constexpr int BLKV = ns_block<NSV>();                                                          \
        const size_t smem = (size_t)2 * NSV * 16 * BLKV;                                               \
        if (PEB) {                                                                                     \
            static bool once = (cudaFuncSetAttribute(KERNEL<GLV, NSV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), true); \
            (void)once;                                                                                \
            KERNEL<GLV, NSV, true><<<ns_grid(GRID, NSV, BLKV), BLKV, smem, ST>>>(__VA_ARGS__);        \
        } else {                                                                                       \
            static bool once = (cudaFuncSetAttribute(KERNEL<GLV, NSV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), true); \
            (void)once;                                                                                \
            KERNEL<GLV, NSV, false><<<ns_grid(GRID, NSV, BLKV), BLKV, smem, ST>>>(__VA_ARGS__);       \
        }                                                                                              \
    }

#define GVA_DISPATCH_NS(GLVAR, NSV, PEB, KERNEL, GRID, ST, ...)                                         \
    switch (GLVAR) {                                                                                   \
        case 1: GVA_DISPATCH_NS2(1, NSV, PEB, KERNEL, GRID, ST, __VA_ARGS__) break;                    \
        case 2: GVA_DISPATCH_NS2(2, NSV, PEB, KERNEL, GRID, ST, __VA_ARGS__) break;                    \
        default: GVA_DISPATCH_NS2(4, NSV, PEB, KERNEL, GRID, ST, __VA_ARGS__) break;                   \
    }

#define GVA_DISPATCH(GLVAR, KERNEL, GRID, ST, ...)                                     \
    switch (GLVAR) {                                                                   \
        case 1: KERNEL<1><<<GRID, kGvaBlock, 0, ST>>>(__VA_ARGS__); break;             \
        case 2: KERNEL<2><<<GRID, kGvaBlock, 0, ST>>>(__VA_ARGS__); break;             \
        default: KERNEL<4><<<GRID, kGvaBlock, 0, ST>>>(__VA_ARGS__); break;            \
    }

static int gva_check(int n, int nsample, int c, int g) {
    if (n < 0 || nsample < 1 || c < 1 || g < 1 || c % g != 0) return AOPT_ERR_INVALID_ARGUMENT;
    return AOPT_OK;
}

extern "C" int aopt_gva_forward(int n, int nsample, int c, int g, const float *value,
                                const float *peb, const float *logits, const int *idx, float *out,
                                float *prob, aopt_stream_t stream) {
    int rc = gva_check(n, nsample, c, g);
    if (rc != AOPT_OK) return rc;
    if (n == 0) return AOPT_OK;
    if (!value || !logits || !idx || !out) return AOPT_ERR_INVALID_ARGUMENT;
    const int I = c / g;
    const int gl = pick_gl(c, I, {value, peb, out});
    if (gl > 0) {
        const long long items = (long long)n * (c / 4);
        const int grid = stride_grid(items, kGvaBlock, 8);
        const bool ns_ok = aligned16(idx);  // the specialised kernels read the idx row as int4
        const int pf = tuning(kTuneL2Prefetch) != 2;
        if (ns_ok && nsample == 16) {
            GVA_DISPATCH_NS(gl, 16, peb != nullptr, gva_forward_ns_kernel, items, as_stream(stream), (long long)n, c, g, value, peb,
                            logits, idx, out, prob, pf);
        } else if (ns_ok && nsample == 8) {
            GVA_DISPATCH_NS(gl, 8, peb != nullptr, gva_forward_ns_kernel, items, as_stream(stream), (long long)n, c, g, value, peb,
                            logits, idx, out, prob, pf);
        } else if (ns_ok && nsample == 32) {
            GVA_DISPATCH_NS(gl, 32, peb != nullptr, gva_forward_ns_kernel, items, as_stream(stream), (long long)n, c, g, value, peb,
                            logits, idx, out, prob, pf);
        } else {
            GVA_DISPATCH(gl, gva_forward_kernel, grid, as_stream(stream), (long long)n, nsample, c, g, value, peb,
                         logits, idx, out, prob);
        }
    } else {
        if (I > kMaxScalarI) return AOPT_ERR_UNSUPPORTED;
        gva_forward_scalar_kernel<<<stride_grid((long long)n * g, kGvaBlock, 8), kGvaBlock, 0, as_stream(stream)>>>(
            n, nsample, c, g, I, value, peb, logits, idx, out, prob);
    }
    return check_launch();
}

extern "C" int aopt_gva_backward_query(int n, int nsample, int c, int g, const float *grad_out,
                                       const float *value, const float *peb, const float *prob,
                                       const int *idx, float *grad_peb, float *grad_logits,
                                       aopt_stream_t stream) {
    int rc = gva_check(n, nsample, c, g);
    if (rc != AOPT_OK) return rc;
    if (n == 0) return AOPT_OK;
    if (!grad_out || !value || !prob || !idx || !grad_logits) return AOPT_ERR_INVALID_ARGUMENT;
    const int I = c / g;
    const int gl = pick_gl(c, I, {grad_out, value, peb, grad_peb});
    if (gl > 0) {
        const long long items = (long long)n * (c / 4);
        const int grid = stride_grid(items, kGvaBlock, 8);
        const bool ns_ok = aligned16(idx);
        if (ns_ok && nsample == 16) {
            GVA_DISPATCH_NS(gl, 16, peb != nullptr, gva_backward_query_ns_kernel, items, as_stream(stream), (long long)n, c, g, grad_out,
                            value, peb, prob, idx, grad_peb, grad_logits);
        } else if (ns_ok && nsample == 8) {
            GVA_DISPATCH_NS(gl, 8, peb != nullptr, gva_backward_query_ns_kernel, items, as_stream(stream), (long long)n, c, g, grad_out,
                            value, peb, prob, idx, grad_peb, grad_logits);
        } else if (ns_ok && nsample == 32) {
            GVA_DISPATCH_NS(gl, 32, peb != nullptr, gva_backward_query_ns_kernel, items, as_stream(stream), (long long)n, c, g, grad_out,
                            value, peb, prob, idx, grad_peb, grad_logits);
        } else {
            GVA_DISPATCH(gl, gva_backward_query_kernel, grid, as_stream(stream), (long long)n, nsample, c, g, grad_out,
                         value, peb, prob, idx, grad_peb, grad_logits);
        }
    } else {
        if (I > kMaxScalarI) return AOPT_ERR_UNSUPPORTED;
        gva_backward_query_scalar_kernel<<<stride_grid((long long)n * g, kGvaBlock, 8), kGvaBlock, 0, as_stream(stream)>>>(
            n, nsample, c, g, I, grad_out, value, peb, prob, idx, grad_peb, grad_logits);
    }
    return check_launch();
}

extern "C" int aopt_gva_backward_value(int n_src, int nsample, int c, int g, const float *grad_out,
                                       const float *prob, const int *rowptr, const int *perm,
                                       float *grad_value, aopt_stream_t stream) {
    int rc = gva_check(n_src, nsample, c, g);
    if (rc != AOPT_OK) return rc;
    if (n_src == 0) return AOPT_OK;
    if (!grad_out || !prob || !rowptr || !perm || !grad_value) return AOPT_ERR_INVALID_ARGUMENT;
    const int I = c / g;
    const int gl = pick_gl(c, I, {grad_out, grad_value});
    if (gl == 2 && bv_group_mode() == 1 && use_batched_walk()) {
        launch_bv_group_walk(n_src, nsample, c, g, grad_out, prob, rowptr, perm, grad_value, as_stream(stream));
        return check_launch();
    }
    if (gl > 0 && use_batched_walk()) {
        if (gl == 1) launch_bv_walk<1>(n_src, nsample, c, g, grad_out, prob, rowptr, perm, grad_value, as_stream(stream));
        else if (gl == 2) launch_bv_walk<2>(n_src, nsample, c, g, grad_out, prob, rowptr, perm, grad_value, as_stream(stream));
        else launch_bv_walk<4>(n_src, nsample, c, g, grad_out, prob, rowptr, perm, grad_value, as_stream(stream));
        return check_launch();
    }
    if (gl > 0) {
        // Measured at L0: unroll4 166 us, batch8 + cp.async 221 us (the row walk is bound by L2 gather
        // bandwidth — ~1.1 GB of grad_out rows per launch — not by latency).  AOPT_BV_IMPL=batch opts in.
        static const bool use_unroll4 = [] { const char *e = getenv("AOPT_BV_IMPL"); return !(e && e[0] == 'b'); }();
        if (use_unroll4) {
            const int grid4 = stride_grid((long long)n_src * (c / 4), kGvaBlock, 8);
            GVA_DISPATCH(gl, gva_backward_value_unroll4_kernel, grid4, as_stream(stream), (long long)n_src, nsample,
                         log2_exact(nsample), c, g, grad_out, prob, rowptr, perm, grad_value);
            return check_launch();
        }
        const int grid = stride_grid((long long)n_src * (c / 4), kBvBlock, 6);
        GVA_DISPATCH(gl, gva_backward_value_kernel, grid, as_stream(stream), (long long)n_src, nsample,
                     log2_exact(nsample), c, g, grad_out, prob, rowptr, perm, grad_value);
    } else {
        if (I > kMaxScalarI) return AOPT_ERR_UNSUPPORTED;
        gva_backward_value_scalar_kernel<<<stride_grid((long long)n_src * g, kGvaBlock, 8), kGvaBlock, 0, as_stream(stream)>>>(
            n_src, nsample, c, g, I, grad_out, prob, rowptr, perm, grad_value);
    }
    return check_launch();
}

// Fused backward for self-attention (n_src == n): grad_peb, grad_logits and grad_value in ONE kernel when the
// specialised path applies (128-bit layout, nsample in {8,16,32}); otherwise the two kernels above back to back.
// AOPT_GVA_BWD=split forces the two-kernel path (A/B measurements).
extern "C" int aopt_gva_backward(int n, int nsample, int c, int g, const float *grad_out, const float *value,
                                 const float *peb, const float *prob, const int *idx, const int *rowptr,
                                 const int *perm, float *grad_peb, float *grad_logits, float *grad_value,
                                 aopt_stream_t stream) {
    int rc = gva_check(n, nsample, c, g);
    if (rc != AOPT_OK) return rc;
    if (n == 0) return AOPT_OK;
    if (!grad_out || !value || !prob || !idx || !grad_logits || !rowptr || !perm || !grad_value)
        return AOPT_ERR_INVALID_ARGUMENT;
    const bool split = tuning(kTuneGvaBwd) == 2;
    const int I = c / g;
    const int gl = pick_gl(c, I, {grad_out, value, peb, grad_peb, grad_value});
    const bool ns = nsample == 8 || nsample == 16 || nsample == 32;
    if (split || gl == 0 || !ns || !aligned16(idx)) {
        rc = aopt_gva_backward_query(n, nsample, c, g, grad_out, value, peb, prob, idx, grad_peb, grad_logits, stream);
        if (rc != AOPT_OK) return rc;
        return aopt_gva_backward_value(n, nsample, c, g, grad_out, prob, rowptr, perm, grad_value, stream);
    }
    const long long items = (long long)n * (c / 4);
    const int pf = tuning(kTuneL2Prefetch) != 2;
    if (nsample == 16) {
        GVA_DISPATCH_NS(gl, 16, peb != nullptr, gva_backward_fused_ns_kernel, items, as_stream(stream), (long long)n, c, g,
                        grad_out, value, peb, prob, idx, rowptr, perm, grad_peb, grad_logits, grad_value, pf);
    } else if (nsample == 8) {
        GVA_DISPATCH_NS(gl, 8, peb != nullptr, gva_backward_fused_ns_kernel, items, as_stream(stream), (long long)n, c, g,
                        grad_out, value, peb, prob, idx, rowptr, perm, grad_peb, grad_logits, grad_value, pf);
    } else {
        GVA_DISPATCH_NS(gl, 32, peb != nullptr, gva_backward_fused_ns_kernel, items, as_stream(stream), (long long)n, c, g,
                        grad_out, value, peb, prob, idx, rowptr, perm, grad_peb, grad_logits, grad_value, pf);
    }
    return check_launch();
}
