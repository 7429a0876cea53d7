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
// Thread mapping: one thread per (point n, group gi); it owns the I = C/G contiguous channels of
// that group (I = 8 in every PTv2 config → two 128-bit accesses per neighbour row).  Lanes of a warp
// hold consecutive groups, so for a fixed neighbour slot a warp reads whole contiguous (n,s) rows of
// peb (32·I·4 bytes).  peb / grad_peb stream with L1::no_allocate; value / grad_out rows are
// re-used by neighbouring points and go through the read-only path.
//
// Algorithmic bytes: forward 4NC + 4NkC + 4NkG(+4NkG prob) + 4Nk + 4NC;
// backward reads 8NC + 4NkC + 4NkG + 8Nk + 4(N+1), writes 4NkC + 4NkG + 4NC.
#include <math.h>
#include <initializer_list>

#include "common.cuh"

namespace aopt {

constexpr int kGvaBlock = 256;

// Channels of one group held in registers.  I4 > 0: I = 4·I4 channels, 128-bit accesses.
// I4 == 0: runtime I (<= kMaxScalarI), scalar accesses — fallback for unusual widths/alignment.
constexpr int kMaxScalarI = 64;

template <int I4>
struct GroupVec {
    float4 v[I4];
    __device__ __forceinline__ void zero(int) {
#pragma unroll
        for (int i = 0; i < I4; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ void load_gather(const float *p, int) {
#pragma unroll
        for (int i = 0; i < I4; ++i) v[i] = ldg_gather4(p + 4 * i);
    }
    __device__ __forceinline__ void load_stream(const float *p, int) {
#pragma unroll
        for (int i = 0; i < I4; ++i) v[i] = ldg_stream4(p + 4 * i);
    }
    __device__ __forceinline__ void add(const GroupVec &o, int) {
#pragma unroll
        for (int i = 0; i < I4; ++i) { v[i].x += o.v[i].x; v[i].y += o.v[i].y; v[i].z += o.v[i].z; v[i].w += o.v[i].w; }
    }
    __device__ __forceinline__ void fma(const GroupVec &a, float s, int) {
#pragma unroll
        for (int i = 0; i < I4; ++i) {
            v[i].x = fmaf(a.v[i].x, s, v[i].x); v[i].y = fmaf(a.v[i].y, s, v[i].y);
            v[i].z = fmaf(a.v[i].z, s, v[i].z); v[i].w = fmaf(a.v[i].w, s, v[i].w);
        }
    }
    __device__ __forceinline__ float dot(const GroupVec &a, int) const {
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < I4; ++i) {
            d = fmaf(v[i].x, a.v[i].x, d); d = fmaf(v[i].y, a.v[i].y, d);
            d = fmaf(v[i].z, a.v[i].z, d); d = fmaf(v[i].w, a.v[i].w, d);
        }
        return d;
    }
    __device__ __forceinline__ void store_scaled_stream(float *p, float s, int) const {
#pragma unroll
        for (int i = 0; i < I4; ++i)
            stg_stream4(p + 4 * i, make_float4(v[i].x * s, v[i].y * s, v[i].z * s, v[i].w * s));
    }
    __device__ __forceinline__ void store(float *p, int) const {
#pragma unroll
        for (int i = 0; i < I4; ++i) *reinterpret_cast<float4 *>(p + 4 * i) = v[i];
    }
};

template <>
struct GroupVec<0> {
    float v[kMaxScalarI];
    __device__ __forceinline__ void zero(int I) { for (int i = 0; i < I; ++i) v[i] = 0.f; }
    __device__ __forceinline__ void load_gather(const float *p, int I) { for (int i = 0; i < I; ++i) v[i] = __ldg(p + i); }
    __device__ __forceinline__ void load_stream(const float *p, int I) { for (int i = 0; i < I; ++i) v[i] = __ldg(p + i); }
    __device__ __forceinline__ void add(const GroupVec &o, int I) { for (int i = 0; i < I; ++i) v[i] += o.v[i]; }
    __device__ __forceinline__ void fma(const GroupVec &a, float s, int I) { for (int i = 0; i < I; ++i) v[i] = fmaf(a.v[i], s, v[i]); }
    __device__ __forceinline__ float dot(const GroupVec &a, int I) const {
        float d = 0.f;
        for (int i = 0; i < I; ++i) d = fmaf(v[i], a.v[i], d);
        return d;
    }
    __device__ __forceinline__ void store_scaled_stream(float *p, float s, int I) const { for (int i = 0; i < I; ++i) p[i] = v[i] * s; }
    __device__ __forceinline__ void store(float *p, int I) const { for (int i = 0; i < I; ++i) p[i] = v[i]; }
};

// ---- forward -------------------------------------------------------------------------------------
template <int I4>
__global__ void __launch_bounds__(kGvaBlock)
gva_forward_kernel(long long n, int k, int c, int g, int I, const float *__restrict__ value,
                   const float *__restrict__ peb, const float *__restrict__ logits,
                   const int *__restrict__ idx, float *__restrict__ out, float *__restrict__ prob) {
    const long long total = n * g;
    const long long step = (long long)gridDim.x * kGvaBlock;
    for (long long t = (long long)blockIdx.x * kGvaBlock + threadIdx.x; t < total; t += step) {
        const long long pt = t / g;
        const int gi = (int)(t - pt * g);
        const float *lg = logits + (size_t)pt * k * g + gi;
        // softmax over the k neighbour slots of this (point, group): torch.softmax(dim=1) semantics,
        // exp(x - max) / sum
        float mx = -INFINITY;
        for (int s = 0; s < k; ++s) mx = fmaxf(mx, __ldg(lg + (size_t)s * g));
        float sum = 0.f;
        for (int s = 0; s < k; ++s) sum += expf(__ldg(lg + (size_t)s * g) - mx);
        GroupVec<I4> acc;
        acc.zero(I);
        const int *ix = idx + (size_t)pt * k;
        const size_t ch0 = (size_t)gi * I;
#pragma unroll 4
        for (int s = 0; s < k; ++s) {
            const float p = expf(__ldg(lg + (size_t)s * g) - mx) / sum;
            if (prob) prob[((size_t)pt * k + s) * g + gi] = p;
            const int j = __ldg(ix + s);
            if (j >= 0) {  // sign(idx+1) mask: padded slots contribute nothing
                GroupVec<I4> v;
                v.load_gather(value + (size_t)j * c + ch0, I);
                if (peb) {
                    GroupVec<I4> pe;
                    pe.load_stream(peb + ((size_t)pt * k + s) * c + ch0, I);
                    v.add(pe, I);
                }
                acc.fma(v, p, I);
            }
        }
        acc.store(out + (size_t)pt * c + ch0, I);
    }
}

// ---- backward, per query: grad_peb and grad_logits ---------------------------------------------
template <int I4>
__global__ void __launch_bounds__(kGvaBlock)
gva_backward_query_kernel(long long n, int k, int c, int g, int I, const float *__restrict__ grad_out,
                          const float *__restrict__ value, const float *__restrict__ peb,
                          const float *__restrict__ prob, const int *__restrict__ idx,
                          float *__restrict__ grad_peb, float *grad_logits) {
    const long long total = n * g;
    const long long step = (long long)gridDim.x * kGvaBlock;
    for (long long t = (long long)blockIdx.x * kGvaBlock + threadIdx.x; t < total; t += step) {
        const long long pt = t / g;
        const int gi = (int)(t - pt * g);
        const size_t ch0 = (size_t)gi * I;
        GroupVec<I4> go;
        go.load_gather(grad_out + (size_t)pt * c + ch0, I);
        const int *ix = idx + (size_t)pt * k;
        const float *pr = prob + (size_t)pt * k * g + gi;
        float *gl = grad_logits + (size_t)pt * k * g + gi;
        float dot = 0.f;  // sum_s p_s * dL/dp_s
#pragma unroll 4
        for (int s = 0; s < k; ++s) {
            const float p = __ldg(pr + (size_t)s * g);
            const int j = __ldg(ix + s);
            float gw = 0.f;  // dL/dp_s = mask_s * <grad_out, value[idx]+peb>
            if (j >= 0) {
                GroupVec<I4> v;
                v.load_gather(value + (size_t)j * c + ch0, I);
                if (peb) {
                    GroupVec<I4> pe;
                    pe.load_stream(peb + ((size_t)pt * k + s) * c + ch0, I);
                    v.add(pe, I);
                }
                gw = go.dot(v, I);
            }
            if (grad_peb) go.store_scaled_stream(grad_peb + ((size_t)pt * k + s) * c + ch0, j >= 0 ? p : 0.f, I);
            dot = fmaf(p, gw, dot);
            gl[(size_t)s * g] = gw;  // parked; finalised below by the same thread
        }
        for (int s = 0; s < k; ++s) {
            const float p = __ldg(pr + (size_t)s * g);
            const float gw = gl[(size_t)s * g];
            gl[(size_t)s * g] = p * (gw - dot);  // softmax backward
        }
    }
}

// ---- backward, per source: grad_value through the CSR -------------------------------------------
template <int I4>
__global__ void __launch_bounds__(kGvaBlock)
gva_backward_value_kernel(long long n_src, int k, int c, int g, int I, const float *__restrict__ grad_out,
                          const float *__restrict__ prob, const int *__restrict__ rowptr,
                          const int *__restrict__ perm, float *__restrict__ grad_value) {
    const long long total = n_src * g;
    const long long step = (long long)gridDim.x * kGvaBlock;
    for (long long t = (long long)blockIdx.x * kGvaBlock + threadIdx.x; t < total; t += step) {
        const long long j = t / g;
        const int gi = (int)(t - j * g);
        const size_t ch0 = (size_t)gi * I;
        GroupVec<I4> acc;
        acc.zero(I);
        const int e_end = __ldg(rowptr + j + 1);
#pragma unroll 4
        for (int e = __ldg(rowptr + j); e < e_end; ++e) {
            const int p = __ldg(perm + e);  // flat (query, slot) position; idx[p] == j >= 0 so mask = 1
            const int q = p / k;
            const float w = __ldg(prob + (size_t)p * g + gi);
            GroupVec<I4> go;
            go.load_gather(grad_out + (size_t)q * c + ch0, I);
            acc.fma(go, w, I);
        }
        acc.store(grad_value + (size_t)j * c + ch0, I);
    }
}

// I4 to use: 128-bit path needs I % 4 == 0, I4 in {1,2,4} and 16-byte aligned pointers.
static int pick_i4(int I, std::initializer_list<const void *> ptrs) {
    if (I % 4 != 0) return 0;
    int i4 = I / 4;
    if (i4 != 1 && i4 != 2 && i4 != 4) return 0;
    for (const void *p : ptrs)
        if (p && !aligned16(p)) return 0;
    return i4;
}

}  // namespace aopt

using namespace aopt;

#define GVA_DISPATCH(I4VAR, KERNEL, GRID, ST, ...)                                     \
    switch (I4VAR) {                                                                   \
        case 1: KERNEL<1><<<GRID, kGvaBlock, 0, ST>>>(__VA_ARGS__); break;             \
        case 2: KERNEL<2><<<GRID, kGvaBlock, 0, ST>>>(__VA_ARGS__); break;             \
        case 4: KERNEL<4><<<GRID, kGvaBlock, 0, ST>>>(__VA_ARGS__); break;             \
        default: KERNEL<0><<<GRID, kGvaBlock, 0, ST>>>(__VA_ARGS__); break;            \
    }

static int gva_check(int n, int nsample, int c, int g) {
    if (n < 0 || nsample < 1 || c < 1 || g < 1 || c % g != 0) return AOPT_ERR_INVALID_ARGUMENT;
    int I = c / g;
    if (I % 4 != 0 || (I != 4 && I != 8 && I != 16)) {
        if (I > kMaxScalarI) return AOPT_ERR_UNSUPPORTED;
    }
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
    const int i4 = pick_i4(I, {value, peb, out});
    const int grid = stride_grid((long long)n * g, kGvaBlock, 8);
    GVA_DISPATCH(i4, gva_forward_kernel, grid, as_stream(stream), (long long)n, nsample, c, g, I, value, peb,
                 logits, idx, out, prob);
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
    const int i4 = pick_i4(I, {grad_out, value, peb, grad_peb});
    const int grid = stride_grid((long long)n * g, kGvaBlock, 8);
    GVA_DISPATCH(i4, gva_backward_query_kernel, grid, as_stream(stream), (long long)n, nsample, c, g, I,
                 grad_out, value, peb, prob, idx, grad_peb, grad_logits);
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
    const int i4 = pick_i4(I, {grad_out, grad_value});
    const int grid = stride_grid((long long)n_src * g, kGvaBlock, 8);
    GVA_DISPATCH(i4, gva_backward_value_kernel, grid, as_stream(stream), (long long)n_src, nsample, c, g, I,
                 grad_out, prob, rowptr, perm, grad_value);
    return check_launch();
}
