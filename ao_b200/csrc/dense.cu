// dense.cu — the BatchNorm-shaped parts of a PTv2 block that sit between the point operators and the dense
// per-point Linear layers ("BN + tiny-K Linear fusion"):
//
//   aopt_bn_act_*        y = [ReLU]( [residual +] [row_scale ·] BatchNorm_train(x) )  on the rows of an (rows, C) matrix
//                        = PointBatchNorm (+ nn.ReLU, + DropPath, + the residual add) of
//                        /root/reference/pointcept/models/point_transformer_v2/point_transformer_v2m2_base.py:25-45,187-197
//                        and of every Linear → PointBatchNorm → ReLU triple (:86-93,240-242,288-295,363-364).
//                        fp32 or bf16 in / out, statistics and arithmetic in fp32 (partial sums combined in fp64).
//   aopt_we_tail_*       logits = Linear(G,G)( ReLU( BatchNorm_train( rel + upe + const ) ) ) on the (rows = N·k, G) tensors:
//                        weight_encoding[1:] of GroupedVectorAttention (:94-99,120) together with the two additions that
//                        form its input in the relation-free schedule (ptv2._forward_fused).  The G x G layer is K = 6 / 12:
//                        CUDA-core FMAs on a row held in registers; its parameter gradients (a (G, rows) x (rows, G) product
//                        with rows = 5.12 M that cuBLAS runs as ONE CTA: 723 us per call, profiles/r02e) are per-thread
//                        accumulators reduced in a fixed order.
//
// All of it is HBM-/L2-bound element-wise and reduction work: one thread keeps one 4-channel column chunk (BN) or one
// row (tail), 128-bit accesses, grids sized in multiples of the SM count, no atomics (per-CTA partials are summed in a
// fixed order by partials_reduce_kernel, so results are bitwise repeatable).  The small kernels of one call are chained
// with programmatic dependent launch.
#include <cuda_bf16.h>

#include "common.cuh"

namespace aopt {

constexpr int kDenseBlock = 256;
constexpr int kDenseCtasPerSm = 4;
constexpr int kDenseMaxGrid = kNumSM * kDenseCtasPerSm;

// ---- typed 4-channel access -------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float4 ld4(const T *p);
template <>
__device__ __forceinline__ float4 ld4<float>(const float *p) {
    return __ldg(reinterpret_cast<const float4 *>(p));
}
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16 *p) {
    const uint2 r = __ldg(reinterpret_cast<const uint2 *>(p));
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&r.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T>
__device__ __forceinline__ void st4(T *p, const float4 &v);
template <>
__device__ __forceinline__ void st4<float>(float *p, const float4 &v) {
    *reinterpret_cast<float4 *>(p) = v;
}
template <>
__device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16 *p, const float4 &v) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<const unsigned *>(&a);
    r.y = *reinterpret_cast<const unsigned *>(&b);
    *reinterpret_cast<uint2 *>(p) = r;
}

// ---- reductions ---------------------------------------------------------------------------------------------------
// NV per-thread values -> one row of NV floats per CTA (warp shuffles, then the warps in order).
template <int NV>
__device__ __forceinline__ void block_reduce_store(float (&v)[NV], float *__restrict__ partial_row) {
    __shared__ float sh[kDenseBlock / 32][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        float x = v[j];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        if (lane == 0) sh[warp][j] = x;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < NV; j += kDenseBlock) {
        float s = sh[0][j];
#pragma unroll
        for (int w = 1; w < kDenseBlock / 32; ++w) s += sh[w][j];
        partial_row[j] = s;
    }
}

// out[v] = Σ_p partials[p][v] in fp64, fixed order.  CTA = 32 values x 32 slices of the partial rows: a thread issues
// its <= 19 loads (592 partial rows) back to back — with 8 slices the 74 dependent-latency loads per thread made this
// 3-CTA kernel the longest of the chain (15 us per call, 220 calls per training step: profiles/r02p).
constexpr int kReduceSlices = 32;
__global__ void __launch_bounds__(32 * kReduceSlices)
partials_reduce_kernel(int parts, int width, int stride, const float *__restrict__ partials, double *__restrict__ out,
                       float *__restrict__ out_f32) {
    __shared__ double sh[kReduceSlices][33];
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int v = blockIdx.x * 32 + lane;
    double s = 0.0;
    if (v < width) {
        int p = slice;
        for (; p + 3 * kReduceSlices < parts; p += 4 * kReduceSlices) {
            const float a = partials[(size_t)p * stride + v], b = partials[(size_t)(p + kReduceSlices) * stride + v];
            const float c = partials[(size_t)(p + 2 * kReduceSlices) * stride + v];
            const float d = partials[(size_t)(p + 3 * kReduceSlices) * stride + v];
            s += (double)a;
            s += (double)b;
            s += (double)c;
            s += (double)d;
        }
        for (; p < parts; p += kReduceSlices) s += (double)partials[(size_t)p * stride + v];
    }
    sh[slice][lane] = s;
    __syncthreads();
    if (slice == 0 && v < width) {
        double t = sh[0][lane];
#pragma unroll
        for (int u = 1; u < kReduceSlices; ++u) t += sh[u][lane];
        if (out) out[v] = t;
        if (out_f32) out_f32[v] = (float)t;
    }
}

struct ChanStat {
    float mean, rstd;
};
__device__ __forceinline__ ChanStat stat_from_sums(double s, double q, double inv_rows, float eps, float *var_out = nullptr) {
    const double m = s * inv_rows;
    double var = q * inv_rows - m * m;
    if (var < 0.0) var = 0.0;
    if (var_out) *var_out = (float)var;
    ChanStat r;
    r.mean = (float)m;
    r.rstd = (float)(1.0 / sqrt(var + (double)eps));
    return r;
}

// ---- BatchNorm (+ ReLU, + residual, + per-row scale) on (rows, C), C % 4 == 0 ----------------------------------
// A thread owns V = 8 (C % 8 == 0) or 4 consecutive channels and walks the rows.  These kernels move 30-250 MB per call:
// what they need is bytes in flight, not occupancy.  The first version (one 8-byte load per thread and iteration) ran at
// 17-35 % of the DRAM peak at level 0 (profiles/r02t_ops_L0_ncu_summary.md rows 0-6); here every thread issues the raw
// 16-byte loads of kRowUnroll rows — of all the tensors it reads — before it unpacks the first one.
constexpr int kRowUnroll = 4;

template <typename T, int V>
struct Raw;
template <>
struct Raw<float, 4> {
    float4 a;
    __device__ __forceinline__ void load(const float *p) { a = __ldg(reinterpret_cast<const float4 *>(p)); }
    __device__ __forceinline__ void unpack(float (&v)[4]) const { v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; }
};
template <>
struct Raw<float, 8> {
    float4 a, b;
    __device__ __forceinline__ void load(const float *p) {
        a = __ldg(reinterpret_cast<const float4 *>(p));
        b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    }
    __device__ __forceinline__ void unpack(float (&v)[8]) const {
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
};
template <>
struct Raw<__nv_bfloat16, 4> {
    uint2 r;
    __device__ __forceinline__ void load(const __nv_bfloat16 *p) { r = __ldg(reinterpret_cast<const uint2 *>(p)); }
    __device__ __forceinline__ void unpack(float (&v)[4]) const {
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&r.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&r.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
};
template <>
struct Raw<__nv_bfloat16, 8> {
    uint4 r;
    __device__ __forceinline__ void load(const __nv_bfloat16 *p) { r = __ldg(reinterpret_cast<const uint4 *>(p)); }
    __device__ __forceinline__ void unpack(float (&v)[8]) const {
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&r.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&r.y));
        const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&r.z));
        const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&r.w));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
    }
};
template <int V>
__device__ __forceinline__ void store_vec(float *p, const float (&v)[V]) {
#pragma unroll
    for (int j = 0; j < V; j += 4) *reinterpret_cast<float4 *>(p + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
}
template <int V>
__device__ __forceinline__ void store_vec(__nv_bfloat16 *p, const float (&v)[V]) {
    unsigned w[V / 2];
#pragma unroll
    for (int j = 0; j < V / 2; ++j) {
        const __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        w[j] = *reinterpret_cast<const unsigned *>(&t);
    }
    if constexpr (V == 8) *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    else *reinterpret_cast<uint2 *>(p) = make_uint2(w[0], w[1]);
}

// Threads of a CTA that own the same column chunk (t, t + cols, ...) are summed in that order by the first `cols`
// threads; partial row = [ A (c floats) | B (c floats) ].
template <int V>
__device__ __forceinline__ void column_reduce_store_v(const float (&a)[V], const float (&b)[V], int cols, int c, int col,
                                                      float *__restrict__ partial_row) {
    __shared__ float sh[2 * V][kDenseBlock];
    const int t = threadIdx.x;
#pragma unroll
    for (int j = 0; j < V; ++j) { sh[j][t] = a[j]; sh[V + j][t] = b[j]; }
    __syncthreads();
    if (t < cols) {
        float A[V], B[V];
#pragma unroll
        for (int j = 0; j < V; ++j) { A[j] = sh[j][t]; B[j] = sh[V + j][t]; }
        for (int u = t + cols; u < kDenseBlock; u += cols) {
#pragma unroll
            for (int j = 0; j < V; ++j) { A[j] += sh[j][u]; B[j] += sh[V + j][u]; }
        }
        store_vec<V>(partial_row + V * col, A);
        store_vec<V>(partial_row + c + V * col, B);
    }
}

template <typename XT, int V>
__global__ void __launch_bounds__(kDenseBlock)
bn_partial_kernel(long long rows, int c, const XT *__restrict__ x, long long ldx, float *__restrict__ partials) {
    const int cols = c / V;
    const ColWalk w = col_walk(cols, kDenseBlock);
    pdl_trigger();
    float s[V], q[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { s[j] = 0.f; q[j] = 0.f; }
    const XT *p = x + V * w.col;
    for (long long row = w.row; row < rows; row += kRowUnroll * w.row_step) {
        Raw<XT, V> r[kRowUnroll];
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
            const long long ru = row + u * w.row_step;
            r[u].load(p + (ru < rows ? ru : row) * ldx);
        }
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
            if (row + u * w.row_step < rows) {
                float v[V];
                r[u].unpack(v);
#pragma unroll
                for (int j = 0; j < V; ++j) { s[j] += v[j]; q[j] = fmaf(v[j], v[j], q[j]); }
            }
        }
    }
    column_reduce_store_v<V>(s, q, cols, c, w.col, partials + (size_t)blockIdx.x * 2 * c);
}

template <typename XT, typename OT, int V>
__global__ void __launch_bounds__(kDenseBlock)
bn_apply_kernel(long long rows, int c, const XT *__restrict__ x, long long ldx, const double *__restrict__ sums, double inv_rows,
                float eps, const float *__restrict__ gamma, const float *__restrict__ beta,
                const OT *__restrict__ residual, const float *__restrict__ row_scale, int relu, OT *__restrict__ out,
                float *__restrict__ stats_out, float *__restrict__ running_mean, float *__restrict__ running_var,
                float momentum, float unbias, const float *__restrict__ mean_shift, long long *__restrict__ batches_tracked) {
    const int cols = c / V;
    const ColWalk w = col_walk(cols, kDenseBlock);
    pdl_wait();
    float mean[V], sc[V], sh[V];
    const bool writer = (long long)blockIdx.x * kDenseBlock + threadIdx.x < cols;   // one thread per column chunk
    if (batches_tracked && blockIdx.x == 0 && threadIdx.x == 0) *batches_tracked += 1;   // nn.BatchNorm1d.num_batches_tracked
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const int ch = V * w.col + j;
        float var;
        const ChanStat st = stat_from_sums(sums[ch], sums[c + ch], inv_rows, eps, &var);
        mean[j] = st.mean;
        sc[j] = __ldg(gamma + ch) * st.rstd;
        sh[j] = __ldg(beta + ch);
        if (writer) {
            stats_out[ch] = st.mean;
            stats_out[c + ch] = st.rstd;
            // mean_shift: a bias the caller left out of x because the normalisation removes it (Linear bias in front of
            // a training-mode BatchNorm); only the running mean sees it
            if (running_mean)
                running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (st.mean + (mean_shift ? mean_shift[ch] : 0.f));
            if (running_var) running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * var * unbias;
        }
    }
    const XT *px = x + V * w.col;
    const OT *pr = residual ? residual + V * w.col : nullptr;
    OT *po = out + V * w.col;
    for (long long row = w.row; row < rows; row += kRowUnroll * w.row_step) {
        Raw<XT, V> rx[kRowUnroll];
        Raw<OT, V> rr[kRowUnroll];
        float rs[kRowUnroll];
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
            const long long ru = row + u * w.row_step;
            const long long rc = ru < rows ? ru : row;
            rx[u].load(px + rc * ldx);
            if (pr) rr[u].load(pr + rc * c);
            rs[u] = row_scale ? __ldg(row_scale + rc) : 1.f;
        }
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
            const long long ru = row + u * w.row_step;
            if (ru < rows) {
                float v[V], y[V];
                rx[u].unpack(v);
#pragma unroll
                for (int j = 0; j < V; ++j) y[j] = fmaf(v[j] - mean[j], sc[j], sh[j]) * rs[u];
                if (pr) {
                    float e[V];
                    rr[u].unpack(e);
#pragma unroll
                    for (int j = 0; j < V; ++j) y[j] += e[j];
                }
                if (relu) {
#pragma unroll
                    for (int j = 0; j < V; ++j) y[j] = fmaxf(y[j], 0.f);
                }
                store_vec<V>(po + ru * c, y);
            }
        }
    }
}

// dy = grad_out ⊙ [out > 0] (· row_scale);  partial row = [ Σ dy | Σ dy·x̂ ]
template <typename XT, typename OT, int V>
__global__ void __launch_bounds__(kDenseBlock)
bn_bwd_partial_kernel(long long rows, int c, const OT *__restrict__ grad_out, const OT *__restrict__ out,
                      const XT *__restrict__ x, long long ldx, const float *__restrict__ stats,
                      const float *__restrict__ row_scale, float *__restrict__ partials) {
    const int cols = c / V;
    const ColWalk w = col_walk(cols, kDenseBlock);
    pdl_trigger();
    float mean[V], rstd[V], s1[V], s2[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        mean[j] = stats[V * w.col + j];
        rstd[j] = stats[c + V * w.col + j];
        s1[j] = 0.f;
        s2[j] = 0.f;
    }
    const XT *px = x + V * w.col;
    const OT *pg = grad_out + V * w.col;
    const OT *po = out ? out + V * w.col : nullptr;
    for (long long row = w.row; row < rows; row += kRowUnroll * w.row_step) {
        Raw<XT, V> rx[kRowUnroll];
        Raw<OT, V> rg[kRowUnroll], ro[kRowUnroll];
        float rs[kRowUnroll];
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
            const long long ru = row + u * w.row_step;
            const long long rc = ru < rows ? ru : row;
            rg[u].load(pg + rc * c);
            rx[u].load(px + rc * ldx);
            if (po) ro[u].load(po + rc * c);
            rs[u] = row_scale ? __ldg(row_scale + rc) : 1.f;
        }
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
            if (row + u * w.row_step < rows) {
                float g[V], v[V];
                rg[u].unpack(g);
                rx[u].unpack(v);
                if (po) {
                    float o[V];
                    ro[u].unpack(o);
#pragma unroll
                    for (int j = 0; j < V; ++j) g[j] = o[j] > 0.f ? g[j] : 0.f;
                }
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const float d = g[j] * rs[u];
                    s1[j] += d;
                    s2[j] = fmaf(d, (v[j] - mean[j]) * rstd[j], s2[j]);
                }
            }
        }
    }
    column_reduce_store_v<V>(s1, s2, cols, c, w.col, partials + (size_t)blockIdx.x * 2 * c);
}

// dx = γ·rstd·(dy − mean(dy) − x̂·mean(dy·x̂));  grad_residual = grad_out ⊙ [out > 0];  dγ = Σ dy·x̂, dβ = Σ dy
template <typename XT, typename OT, int V>
__global__ void __launch_bounds__(kDenseBlock)
bn_bwd_apply_kernel(long long rows, int c, const OT *__restrict__ grad_out, const OT *__restrict__ out,
                    const XT *__restrict__ x, long long ldx, const float *__restrict__ stats, const float *__restrict__ gamma,
                    const float *__restrict__ row_scale, const double *__restrict__ sums, double inv_rows,
                    XT *__restrict__ grad_x, long long ldgx, OT *__restrict__ grad_residual, float *__restrict__ grad_gamma,
                    float *__restrict__ grad_beta) {
    const int cols = c / V;
    const ColWalk w = col_walk(cols, kDenseBlock);
    pdl_wait();
    const bool writer = (long long)blockIdx.x * kDenseBlock + threadIdx.x < cols;
    float mean[V], rstd[V], a[V], m1[V], m2[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const int ch = V * w.col + j;
        mean[j] = stats[ch];
        rstd[j] = stats[c + ch];
        a[j] = __ldg(gamma + ch) * rstd[j];
        const double S1 = sums[ch], S2 = sums[c + ch];
        m1[j] = (float)(S1 * inv_rows);
        m2[j] = (float)(S2 * inv_rows);
        if (writer) {
            grad_beta[ch] = (float)S1;
            grad_gamma[ch] = (float)S2;
        }
    }
    const XT *px = x + V * w.col;
    const OT *pg = grad_out + V * w.col;
    const OT *po = out ? out + V * w.col : nullptr;
    XT *pdx = grad_x + V * w.col;
    OT *pdr = grad_residual ? grad_residual + V * w.col : nullptr;
    for (long long row = w.row; row < rows; row += kRowUnroll * w.row_step) {
        Raw<XT, V> rx[kRowUnroll];
        Raw<OT, V> rg[kRowUnroll], ro[kRowUnroll];
        float rs[kRowUnroll];
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
            const long long ru = row + u * w.row_step;
            const long long rc = ru < rows ? ru : row;
            rg[u].load(pg + rc * c);
            rx[u].load(px + rc * ldx);
            if (po) ro[u].load(po + rc * c);
            rs[u] = row_scale ? __ldg(row_scale + rc) : 1.f;
        }
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
            const long long ru = row + u * w.row_step;
            if (ru < rows) {
                float g[V], v[V], d[V];
                rg[u].unpack(g);
                rx[u].unpack(v);
                if (po) {
                    float o[V];
                    ro[u].unpack(o);
#pragma unroll
                    for (int j = 0; j < V; ++j) g[j] = o[j] > 0.f ? g[j] : 0.f;
                }
                if (pdr) store_vec<V>(pdr + ru * c, g);
#pragma unroll
                for (int j = 0; j < V; ++j) d[j] = a[j] * (g[j] * rs[u] - m1[j] - (v[j] - mean[j]) * rstd[j] * m2[j]);
                store_vec<V>(pdx + ru * ldgx, d);
            }
        }
    }
}

// ---- weight-encoding tail on (rows, G): u = rel + upe + cst;  logits = W2 · ReLU(BN(u)) + b2 --------------------------
// First addend of u: either a materialised (rows, G) tensor `rel`, or — gather mode — key / query projections
// kp, qp (N, G) and the neighbour list: rel[row] = kp[idx[row]] - qp[row / k] (key row = 0 where idx < 0), the G-wide
// gva_relation formed while loading, so that the (N, k, G) tensor is neither written nor read twice.
struct TailSrc {
    const float *rel, *kp, *qp;
    const int *idx;
    int k;
};

template <int G>
struct TailRow {
    float u[G];
    __device__ __forceinline__ void load(const TailSrc &src, const float *__restrict__ upe, const float *cst, long long row) {
        if (src.rel) {
            const float2 *a = reinterpret_cast<const float2 *>(src.rel + row * G);
#pragma unroll
            for (int j = 0; j < G / 2; ++j) {
                const float2 v = __ldg(a + j);
                u[2 * j] = v.x;
                u[2 * j + 1] = v.y;
            }
        } else {
            const int jn = __ldg(src.idx + row);
            const float2 *q = reinterpret_cast<const float2 *>(src.qp + (row / src.k) * G);
            const float2 *kk = reinterpret_cast<const float2 *>(src.kp + (size_t)max(jn, 0) * G);
#pragma unroll
            for (int j = 0; j < G / 2; ++j) {
                const float2 a = __ldg(kk + j), b = __ldg(q + j);
                u[2 * j] = (jn >= 0 ? a.x : 0.f) - b.x;
                u[2 * j + 1] = (jn >= 0 ? a.y : 0.f) - b.y;
            }
        }
        if (upe) {
            const float2 *b = reinterpret_cast<const float2 *>(upe + row * G);
#pragma unroll
            for (int j = 0; j < G / 2; ++j) {
                const float2 v = __ldg(b + j);
                u[2 * j] += v.x;
                u[2 * j + 1] += v.y;
            }
        }
        if (cst) {
#pragma unroll
            for (int j = 0; j < G; ++j) u[j] += cst[j];
        }
    }
};

template <int G>
__global__ void __launch_bounds__(kDenseBlock)
we_partial_kernel(long long rows, TailSrc src, const float *__restrict__ upe,
                  const float *__restrict__ cst, float *__restrict__ partials) {
    __shared__ float s_c[G];
    pdl_trigger();
    if (cst && threadIdx.x < G) s_c[threadIdx.x] = cst[threadIdx.x];
    __syncthreads();
    float acc[2 * G];
#pragma unroll
    for (int j = 0; j < 2 * G; ++j) acc[j] = 0.f;
    for (long long row = (long long)blockIdx.x * kDenseBlock + threadIdx.x; row < rows;
         row += (long long)gridDim.x * kDenseBlock) {
        TailRow<G> r;
        r.load(src, upe, cst ? s_c : nullptr, row);
#pragma unroll
        for (int j = 0; j < G; ++j) {
            acc[j] += r.u[j];
            acc[G + j] = fmaf(r.u[j], r.u[j], acc[G + j]);
        }
    }
    block_reduce_store<2 * G>(acc, partials + (size_t)blockIdx.x * 2 * G);
}

template <int G>
struct TailParams {   // shared-memory copy of the per-channel maps and the G x G layer
    float cst[G], mean[G], sc[G], beta[G], rstd[G], w[G * G], b[G];
};

template <int G>
__global__ void __launch_bounds__(kDenseBlock)
we_apply_kernel(long long rows, TailSrc src, const float *__restrict__ upe,
                const float *__restrict__ cst, const double *__restrict__ sums, double inv_rows, float eps,
                const float *__restrict__ gamma, const float *__restrict__ beta, const float *__restrict__ w2,
                const float *__restrict__ b2, float *__restrict__ logits, float *__restrict__ stats_out,
                float *__restrict__ running_mean, float *__restrict__ running_var, float momentum, float unbias,
                long long *__restrict__ batches_tracked) {
    __shared__ TailParams<G> P;
    pdl_wait();
    const int t = threadIdx.x;
    if (batches_tracked && blockIdx.x == 0 && t == 0) *batches_tracked += 1;
    if (t < G) {
        float var;
        const ChanStat st = stat_from_sums(sums[t], sums[G + t], inv_rows, eps, &var);
        P.cst[t] = cst ? cst[t] : 0.f;
        P.mean[t] = st.mean;
        P.sc[t] = gamma[t] * st.rstd;
        P.beta[t] = beta[t];
        P.b[t] = b2 ? b2[t] : 0.f;
        if (blockIdx.x == 0) {
            stats_out[t] = st.mean;
            stats_out[G + t] = st.rstd;
            if (running_mean) running_mean[t] = (1.f - momentum) * running_mean[t] + momentum * st.mean;
            if (running_var) running_var[t] = (1.f - momentum) * running_var[t] + momentum * var * unbias;
        }
    }
    for (int i = t; i < G * G; i += kDenseBlock) P.w[i] = w2[i];
    __syncthreads();
    for (long long row = (long long)blockIdx.x * kDenseBlock + t; row < rows; row += (long long)gridDim.x * kDenseBlock) {
        TailRow<G> r;
        r.load(src, upe, cst ? P.cst : nullptr, row);
        float h[G];
#pragma unroll
        for (int j = 0; j < G; ++j) h[j] = fmaxf(fmaf(r.u[j] - P.mean[j], P.sc[j], P.beta[j]), 0.f);
        float2 *o = reinterpret_cast<float2 *>(logits + row * G);
#pragma unroll
        for (int i = 0; i < G; i += 2) {
            float y0 = P.b[i], y1 = P.b[i + 1];
#pragma unroll
            for (int j = 0; j < G; ++j) {
                y0 = fmaf(P.w[i * G + j], h[j], y0);
                y1 = fmaf(P.w[(i + 1) * G + j], h[j], y1);
            }
            o[i / 2] = make_float2(y0, y1);
        }
    }
}

// Backward, pass 1.  Per row: h (recomputed), dl = grad_logits row, dh = (W2ᵀ dl) ⊙ [h > 0].
// partial row (width 3G + G·G) = [ Σ dh | Σ dh·x̂ | Σ dl | Σ dl_i·h_j ].  gridDim.y slices of IB rows of the G x G
// block keep the per-thread accumulators in registers; slice 0 also owns the three G-wide sums.
template <int G, int IB>
__global__ void __launch_bounds__(kDenseBlock)
we_bwd_partial_kernel(long long rows, TailSrc src, const float *__restrict__ upe,
                      const float *__restrict__ cst, const float *__restrict__ grad_logits,
                      const float *__restrict__ stats, const float *__restrict__ gamma, const float *__restrict__ beta,
                      const float *__restrict__ w2, float *__restrict__ partials) {
    __shared__ TailParams<G> P;
    pdl_trigger();
    const int t = threadIdx.x;
    if (t < G) {
        P.cst[t] = cst ? cst[t] : 0.f;
        P.mean[t] = stats[t];
        P.rstd[t] = stats[G + t];
        P.sc[t] = gamma[t] * stats[G + t];
        P.beta[t] = beta[t];
    }
    for (int i = t; i < G * G; i += kDenseBlock) P.w[i] = w2[i];
    __syncthreads();
    const int i0 = blockIdx.y * IB;
    const bool lead = blockIdx.y == 0;
    float s[3 * G], wacc[IB * G];
#pragma unroll
    for (int j = 0; j < 3 * G; ++j) s[j] = 0.f;
#pragma unroll
    for (int j = 0; j < IB * G; ++j) wacc[j] = 0.f;
    for (long long row = (long long)blockIdx.x * kDenseBlock + t; row < rows; row += (long long)gridDim.x * kDenseBlock) {
        TailRow<G> r;
        r.load(src, upe, cst ? P.cst : nullptr, row);
        float dl[G], h[G];
        const float2 *gp = reinterpret_cast<const float2 *>(grad_logits + row * G);
#pragma unroll
        for (int j = 0; j < G / 2; ++j) {
            const float2 v = __ldg(gp + j);
            dl[2 * j] = v.x;
            dl[2 * j + 1] = v.y;
        }
#pragma unroll
        for (int j = 0; j < G; ++j) h[j] = fmaxf(fmaf(r.u[j] - P.mean[j], P.sc[j], P.beta[j]), 0.f);
        // Σ dl_i·h_j for this slice's rows i0 .. i0 + IB - 1 (i0 is CTA-uniform: the select compiles to a shared index)
#pragma unroll
        for (int i = 0; i < IB; ++i) {
            float di = 0.f;
#pragma unroll
            for (int q = 0; q < G; ++q) di = (q == i0 + i) ? dl[q] : di;
#pragma unroll
            for (int j = 0; j < G; ++j) wacc[i * G + j] = fmaf(di, h[j], wacc[i * G + j]);
        }
        if (lead) {
#pragma unroll
            for (int j = 0; j < G; ++j) {
                float dh = 0.f;
#pragma unroll
                for (int i = 0; i < G; ++i) dh = fmaf(P.w[i * G + j], dl[i], dh);
                dh = h[j] > 0.f ? dh : 0.f;
                s[j] += dh;
                s[G + j] = fmaf(dh, (r.u[j] - P.mean[j]) * P.rstd[j], s[G + j]);
                s[2 * G + j] += dl[j];
            }
        }
    }
    float *prow = partials + (size_t)blockIdx.x * (3 * G + G * G);
    if (lead) block_reduce_store<3 * G>(s, prow);
    __syncthreads();
    block_reduce_store<IB * G>(wacc, prow + 3 * G + i0 * G);
}

// Backward, pass 2: du = γ·rstd·(dh − mean(dh) − x̂·mean(dh·x̂)); parameter gradients from the reduced sums.
template <int G>
__global__ void __launch_bounds__(kDenseBlock)
we_bwd_apply_kernel(long long rows, TailSrc src, const float *__restrict__ upe,
                    const float *__restrict__ cst, const float *__restrict__ grad_logits,
                    const float *__restrict__ stats, const float *__restrict__ gamma, const float *__restrict__ beta,
                    const float *__restrict__ w2, const double *__restrict__ sums, double inv_rows,
                    float *__restrict__ grad_u, float *__restrict__ grad_gamma, float *__restrict__ grad_beta,
                    float *__restrict__ grad_b2, float *__restrict__ grad_w2) {
    __shared__ TailParams<G> P;
    __shared__ float m1[G], m2[G];
    pdl_wait();
    const int t = threadIdx.x;
    if (t < G) {
        P.cst[t] = cst ? cst[t] : 0.f;
        P.mean[t] = stats[t];
        P.rstd[t] = stats[G + t];
        P.sc[t] = gamma[t] * stats[G + t];
        P.beta[t] = beta[t];
        m1[t] = (float)(sums[t] * inv_rows);
        m2[t] = (float)(sums[G + t] * inv_rows);
        if (blockIdx.x == 0) {
            grad_beta[t] = (float)sums[t];
            grad_gamma[t] = (float)sums[G + t];
            grad_b2[t] = (float)sums[2 * G + t];
        }
    }
    for (int i = t; i < G * G; i += kDenseBlock) {
        P.w[i] = w2[i];
        if (blockIdx.x == 0) grad_w2[i] = (float)sums[3 * G + i];
    }
    __syncthreads();
    for (long long row = (long long)blockIdx.x * kDenseBlock + t; row < rows; row += (long long)gridDim.x * kDenseBlock) {
        TailRow<G> r;
        r.load(src, upe, cst ? P.cst : nullptr, row);
        float dl[G];
        const float2 *gp = reinterpret_cast<const float2 *>(grad_logits + row * G);
#pragma unroll
        for (int j = 0; j < G / 2; ++j) {
            const float2 v = __ldg(gp + j);
            dl[2 * j] = v.x;
            dl[2 * j + 1] = v.y;
        }
        float2 *o = reinterpret_cast<float2 *>(grad_u + row * G);
        float du[2];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const float xc = r.u[j] - P.mean[j];
            const float hpre = fmaf(xc, P.sc[j], P.beta[j]);
            float dh = 0.f;
#pragma unroll
            for (int i = 0; i < G; ++i) dh = fmaf(P.w[i * G + j], dl[i], dh);
            dh = hpre > 0.f ? dh : 0.f;
            du[j & 1] = P.sc[j] * (dh - m1[j] - xc * P.rstd[j] * m2[j]);
            if (j & 1) o[j / 2] = make_float2(du[0], du[1]);
        }
    }
}

// ---- small dense helpers of the Linear layers' backward passes ----------------------------------------------------------
// dst[r, j] = src[r, j] (+ bias[j]) for a c-wide column block of two row-major matrices with their own row strides and
// element types: the v block of the fused q|k|v product -> dense fp32 (+ linear_v.bias), and its gradient back into the
// (N, 3C) gradient buffer (one kernel instead of a strided cast + an add).
template <typename ST, typename DT>
__global__ void __launch_bounds__(kDenseBlock)
copy_cols_kernel(long long rows, int c, const ST *__restrict__ src, long long ld_src, const float *__restrict__ bias,
                 DT *__restrict__ dst, long long ld_dst) {
    const int cols = c >> 2;
    const ColWalk w = col_walk(cols, kDenseBlock);
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) b = *reinterpret_cast<const float4 *>(bias + 4 * w.col);
    const ST *ps = src + 4 * w.col;
    DT *pd = dst + 4 * w.col;
    for (long long row = w.row; row < rows; row += w.row_step) {
        float4 v = ld4(ps + row * ld_src);
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        st4(pd + row * ld_dst, v);
    }
}

// One row of G values (fp32 or bf16, dense rows of G elements) into registers; 8-byte loads when rows stay aligned.
template <typename GT, int G>
__device__ __forceinline__ void load_small_row(const GT *__restrict__ g, long long row, float (&gr)[G]) {
    if constexpr (sizeof(GT) == 4) {
        if constexpr ((G & 1) == 0) {
            const float2 *gp = reinterpret_cast<const float2 *>(g + row * G);
#pragma unroll
            for (int i = 0; i < G / 2; ++i) {
                const float2 t = __ldg(gp + i);
                gr[2 * i] = t.x;
                gr[2 * i + 1] = t.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < G; ++i) gr[i] = __ldg(g + row * G + i);
        }
    } else {
        if constexpr ((G & 1) == 0) {
            const __nv_bfloat162 *gp = reinterpret_cast<const __nv_bfloat162 *>(g + row * G);
#pragma unroll
            for (int i = 0; i < G / 2; ++i) {
                const float2 t = __bfloat1622float2(gp[i]);
                gr[2 * i] = t.x;
                gr[2 * i + 1] = t.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < G; ++i) gr[i] = __bfloat162float(g[row * G + i]);
        }
    }
}

// Weight gradient of a Linear with a handful of outputs: out[i, :] = Σ_r g[r, i] · x[r, :], g (rows, G), x (rows, c).
// cuBLAS runs this (G, rows) x (rows, c) product — rows = 320 000, G = 6 — at 170 us (profiles/r02r: a single column of
// CTAs walks the whole K dimension); here a thread keeps a 4-channel column chunk of x, walks the rows and accumulates
// its G x 4 block; per-CTA partials are summed in a fixed order (partials_reduce_kernel).
template <typename GT, typename XT, int G>
__global__ void __launch_bounds__(kDenseBlock)
skinny_wgrad_kernel(long long rows, int c, const GT *__restrict__ g, const XT *__restrict__ x, long long ldx,
                    float *__restrict__ partials) {
    const int cols = c >> 2;
    const ColWalk w = col_walk(cols, kDenseBlock);
    pdl_trigger();
    float4 acc[G];
#pragma unroll
    for (int i = 0; i < G; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const XT *px = x + 4 * w.col;
    for (long long row = w.row; row < rows; row += w.row_step) {
        const float4 v = ld4(px + row * ldx);
        float gr[G];
        load_small_row<GT, G>(g, row, gr);
#pragma unroll
        for (int i = 0; i < G; ++i) {
            acc[i].x = fmaf(gr[i], v.x, acc[i].x);
            acc[i].y = fmaf(gr[i], v.y, acc[i].y);
            acc[i].z = fmaf(gr[i], v.z, acc[i].z);
            acc[i].w = fmaf(gr[i], v.w, acc[i].w);
        }
    }
    // threads of the CTA with the same column chunk, in thread order (as column_reduce_store, G rows at a time)
    __shared__ float4 sh[kDenseBlock];
    float *prow = partials + (size_t)blockIdx.x * G * c;
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < G; ++i) {
        __syncthreads();
        sh[t] = acc[i];
        __syncthreads();
        if (t < cols) {
            float4 A = sh[t];
            for (int u = t + cols; u < kDenseBlock; u += cols) {
                const float4 y = sh[u];
                A.x += y.x; A.y += y.y; A.z += y.z; A.w += y.w;
            }
            *reinterpret_cast<float4 *>(prow + (size_t)i * c + 4 * w.col) = A;
        }
    }
}

// Linear with a handful of outputs, forward and input gradient (the key / query projections by weight_encoding[0] in the
// relation-free schedule: (320 000, 48) x (48, 6)).  cuBLAS serves both shapes — N = 6 and K = 6 — with a 32x32 WMMA
// kernel at ~170 us; they are 40 MB of traffic.
//   out[r, i] = Σ_j x[r, j] · w[i, j]         thread = row, the G x c weights in shared memory (broadcast reads)
//   gx[r, j]  = Σ_i g[r, i] · w[i, j]         thread = (row, 4-channel chunk), its G x 4 weights in registers
template <typename XT, int G>
__global__ void __launch_bounds__(kDenseBlock)
skinny_linear_kernel(long long rows, int c, const XT *__restrict__ x, long long ldx, const float *__restrict__ w,
                     const float *__restrict__ bias, float *__restrict__ out) {
    extern __shared__ float4 s_w[];   // [c / 4][G] : chunk-major so that a thread walks it linearly
    const int cols = c >> 2;
    for (int t = threadIdx.x; t < cols * G; t += kDenseBlock) {
        const int i = t % G, ch = t / G;
        s_w[t] = *reinterpret_cast<const float4 *>(w + (size_t)i * c + 4 * ch);
    }
    __syncthreads();
    for (long long row = (long long)blockIdx.x * kDenseBlock + threadIdx.x; row < rows; row += (long long)gridDim.x * kDenseBlock) {
        float acc[G];
#pragma unroll
        for (int i = 0; i < G; ++i) acc[i] = bias ? __ldg(bias + i) : 0.f;
        const XT *px = x + row * ldx;
        for (int ch = 0; ch < cols; ++ch) {
            const float4 v = ld4(px + 4 * ch);
#pragma unroll
            for (int i = 0; i < G; ++i) {
                const float4 ww = s_w[ch * G + i];
                acc[i] = fmaf(v.x, ww.x, fmaf(v.y, ww.y, fmaf(v.z, ww.z, fmaf(v.w, ww.w, acc[i]))));
            }
        }
        if constexpr ((G & 1) == 0) {
            float2 *o = reinterpret_cast<float2 *>(out + row * G);
#pragma unroll
            for (int i = 0; i < G; i += 2) o[i / 2] = make_float2(acc[i], acc[i + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < G; ++i) out[row * G + i] = acc[i];
        }
    }
}

template <typename XT, int G>
__global__ void __launch_bounds__(kDenseBlock)
skinny_dgrad_kernel(long long rows, int c, const float *__restrict__ g, const float *__restrict__ w,
                    XT *__restrict__ gx, long long ldgx) {
    const int cols = c >> 2;
    const ColWalk cw = col_walk(cols, kDenseBlock);
    float4 ww[G];
#pragma unroll
    for (int i = 0; i < G; ++i) ww[i] = *reinterpret_cast<const float4 *>(w + (size_t)i * c + 4 * cw.col);
    XT *po = gx + 4 * cw.col;
    for (long long row = cw.row; row < rows; row += cw.row_step) {
        float gr[G];
        load_small_row<float, G>(g, row, gr);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < G; ++i) {
            a.x = fmaf(gr[i], ww[i].x, a.x); a.y = fmaf(gr[i], ww[i].y, a.y);
            a.z = fmaf(gr[i], ww[i].z, a.z); a.w = fmaf(gr[i], ww[i].w, a.w);
        }
        st4(po + row * ldgx, a);
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------
static size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }

static int bn_grid(long long rows, int c) { return col_grid(rows, c >> 2, kDenseBlock, kDenseCtasPerSm); }
// the row-unrolled BatchNorm kernels: V channels per thread, kRowUnroll rows per iteration; grids sized to what is
// resident at once (no tail wave): 4 CTAs per SM for the 58-register statistics kernel, 2 for the 92-128-register ones
static int bn_vec(int c) { return (c & 7) == 0 ? 8 : 4; }
static int bn_grid_v(long long rows, int c, int ctas_per_sm = 2) {
    const int cols = c / bn_vec(c);
    return col_grid((rows + kRowUnroll - 1) / kRowUnroll, cols, kDenseBlock, ctas_per_sm);
}
static int row_grid(long long rows) { return stride_grid(rows, kDenseBlock, kDenseCtasPerSm); }
static bool bn_width_ok(int c) { return c >= 4 && (c & 3) == 0 && c <= 4 * kDenseBlock; }

template <typename XT, typename OT>
static int bn_forward_t(long long rows, int c, const void *x, long long ldx, const float *gamma, const float *beta, float eps,
                        const void *residual, const float *row_scale, int relu, void *out, float *stats_out,
                        float *running_mean, float *running_var, float momentum, const float *mean_shift,
                        long long *batches_tracked, float *partials, double *sums, cudaStream_t st) {
    const int grid_p = bn_grid_v(rows, c, 4), grid = bn_grid_v(rows, c, 2);
    const bool pdl = tuning(kTunePdl) != 2;
    const bool v8 = bn_vec(c) == 8;
    if (v8) bn_partial_kernel<XT, 8><<<grid_p, kDenseBlock, 0, st>>>(rows, c, static_cast<const XT *>(x), ldx, partials);
    else bn_partial_kernel<XT, 4><<<grid_p, kDenseBlock, 0, st>>>(rows, c, static_cast<const XT *>(x), ldx, partials);
    launch_chain(pdl, partials_reduce_kernel, div_up(2 * c, 32), 32 * kReduceSlices, 0, st, grid_p, 2 * c, 2 * c, (const float *)partials, sums,
                 (float *)nullptr);
    const float unbias = rows > 1 ? (float)((double)rows / (double)(rows - 1)) : 1.f;
    if (v8)
        launch_chain(pdl, bn_apply_kernel<XT, OT, 8>, grid, kDenseBlock, 0, st, rows, c, static_cast<const XT *>(x), ldx,
                     (const double *)sums, 1.0 / (double)rows, eps, gamma, beta, static_cast<const OT *>(residual), row_scale,
                     relu, static_cast<OT *>(out), stats_out, running_mean, running_var, momentum, unbias, mean_shift,
                     batches_tracked);
    else
        launch_chain(pdl, bn_apply_kernel<XT, OT, 4>, grid, kDenseBlock, 0, st, rows, c, static_cast<const XT *>(x), ldx,
                     (const double *)sums, 1.0 / (double)rows, eps, gamma, beta, static_cast<const OT *>(residual), row_scale,
                     relu, static_cast<OT *>(out), stats_out, running_mean, running_var, momentum, unbias, mean_shift,
                     batches_tracked);
    return check_launch(3);
}

template <typename XT, typename OT>
static int bn_backward_t(long long rows, int c, const void *grad_out, const void *out, const void *x, long long ldx,
                         const float *stats, const float *gamma, const float *row_scale, void *grad_x, long long ldgx,
                         void *grad_residual,
                         float *grad_gamma, float *grad_beta, float *partials, double *sums, cudaStream_t st) {
    const int grid = bn_grid_v(rows, c);
    const bool pdl = tuning(kTunePdl) != 2;
    const bool v8 = bn_vec(c) == 8;
    if (v8)
        bn_bwd_partial_kernel<XT, OT, 8><<<grid, kDenseBlock, 0, st>>>(rows, c, static_cast<const OT *>(grad_out),
                                                                       static_cast<const OT *>(out),
                                                                       static_cast<const XT *>(x), ldx, stats, row_scale, partials);
    else
        bn_bwd_partial_kernel<XT, OT, 4><<<grid, kDenseBlock, 0, st>>>(rows, c, static_cast<const OT *>(grad_out),
                                                                       static_cast<const OT *>(out),
                                                                       static_cast<const XT *>(x), ldx, stats, row_scale, partials);
    launch_chain(pdl, partials_reduce_kernel, div_up(2 * c, 32), 32 * kReduceSlices, 0, st, grid, 2 * c, 2 * c, (const float *)partials, sums,
                 (float *)nullptr);
    if (v8)
        launch_chain(pdl, bn_bwd_apply_kernel<XT, OT, 8>, grid, kDenseBlock, 0, st, rows, c, static_cast<const OT *>(grad_out),
                     static_cast<const OT *>(out), static_cast<const XT *>(x), ldx, stats, gamma, row_scale, (const double *)sums,
                     1.0 / (double)rows, static_cast<XT *>(grad_x), ldgx, static_cast<OT *>(grad_residual), grad_gamma, grad_beta);
    else
        launch_chain(pdl, bn_bwd_apply_kernel<XT, OT, 4>, grid, kDenseBlock, 0, st, rows, c, static_cast<const OT *>(grad_out),
                     static_cast<const OT *>(out), static_cast<const XT *>(x), ldx, stats, gamma, row_scale, (const double *)sums,
                     1.0 / (double)rows, static_cast<XT *>(grad_x), ldgx, static_cast<OT *>(grad_residual), grad_gamma, grad_beta);
    return check_launch(3);
}

template <int G, int IB>
static int we_forward_t(long long rows, TailSrc rel, const float *upe, const float *cst, const float *gamma,
                        const float *beta, float eps, const float *w2, const float *b2, float *logits, float *stats_out,
                        float *running_mean, float *running_var, float momentum, long long *batches_tracked, float *partials,
                        double *sums, cudaStream_t st) {
    const int grid = row_grid(rows);
    const bool pdl = tuning(kTunePdl) != 2;
    we_partial_kernel<G><<<grid, kDenseBlock, 0, st>>>(rows, rel, upe, cst, partials);
    launch_chain(pdl, partials_reduce_kernel, div_up(2 * G, 32), 32 * kReduceSlices, 0, st, grid, 2 * G, 2 * G, (const float *)partials, sums,
                 (float *)nullptr);
    const float unbias = rows > 1 ? (float)((double)rows / (double)(rows - 1)) : 1.f;
    launch_chain(pdl, we_apply_kernel<G>, grid, kDenseBlock, 0, st, rows, rel, upe, cst, (const double *)sums,
                 1.0 / (double)rows, eps, gamma, beta, w2, b2, logits, stats_out, running_mean, running_var, momentum, unbias,
                 batches_tracked);
    return check_launch(3);
}

template <int G, int IB>
static int we_backward_t(long long rows, TailSrc rel, const float *upe, const float *cst, const float *grad_logits,
                         const float *stats, const float *gamma, const float *beta, const float *w2, float *grad_u,
                         float *grad_gamma, float *grad_beta, float *grad_b2, float *grad_w2, float *partials, double *sums,
                         cudaStream_t st) {
    const int grid = row_grid(rows);
    const bool pdl = tuning(kTunePdl) != 2;
    const int width = 3 * G + G * G;
    we_bwd_partial_kernel<G, IB><<<dim3(grid, G / IB), kDenseBlock, 0, st>>>(rows, rel, upe, cst, grad_logits, stats, gamma,
                                                                             beta, w2, partials);
    launch_chain(pdl, partials_reduce_kernel, div_up(width, 32), 32 * kReduceSlices, 0, st, grid, width, width, (const float *)partials, sums,
                 (float *)nullptr);
    launch_chain(pdl, we_bwd_apply_kernel<G>, grid, kDenseBlock, 0, st, rows, rel, upe, cst, grad_logits, stats, gamma, beta,
                 w2, (const double *)sums, 1.0 / (double)rows, grad_u, grad_gamma, grad_beta, grad_b2, grad_w2);
    return check_launch(3);
}

}  // namespace aopt

using namespace aopt;

extern "C" int aopt_bn_act_supported(int c) { return bn_width_ok(c) ? 1 : 0; }
extern "C" int aopt_we_tail_supported(int g) { return (g == 6 || g == 12) ? 1 : 0; }

// partial rows (<= kDenseMaxGrid of them) + the fp64 sums
extern "C" size_t aopt_dense_workspace_bytes(int width) {
    if (width < 1) width = 1;
    return a256(4 * (size_t)kDenseMaxGrid * width) + a256(8 * (size_t)width);
}

static bool carve_dense(void *ws, size_t ws_bytes, int width, float **partials, double **sums) {
    if (!ws || ws_bytes < aopt_dense_workspace_bytes(width)) return false;
    *partials = static_cast<float *>(ws);
    *sums = reinterpret_cast<double *>(static_cast<char *>(ws) + a256(4 * (size_t)kDenseMaxGrid * width));
    return true;
}

extern "C" int aopt_bn_act_forward(int64_t rows, int c, const void *x, int64_t ldx, int x_dtype, const float *gamma, const float *beta,
                                   float eps, const void *residual, const float *row_scale, int relu, void *out,
                                   int out_dtype, float *stats_out, float *running_mean, float *running_var,
                                   float momentum, const float *mean_shift, long long *batches_tracked, void *workspace,
                                   size_t workspace_bytes, aopt_stream_t stream) {
    if (rows <= 0 || !bn_width_ok(c) || !x || !gamma || !beta || !out || !stats_out) return AOPT_ERR_INVALID_ARGUMENT;
    if ((x_dtype | out_dtype) & ~1) return AOPT_ERR_INVALID_ARGUMENT;
    if (ldx < c || (ldx & 3)) return AOPT_ERR_INVALID_ARGUMENT;   // rows of x stay 4-element aligned
    float *partials;
    double *sums;
    if (!carve_dense(workspace, workspace_bytes, 2 * c, &partials, &sums)) return AOPT_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
#define AOPT_BN_FWD(XT, OT)                                                                                               \
    return bn_forward_t<XT, OT>(rows, c, x, ldx, gamma, beta, eps, residual, row_scale, relu, out, stats_out, running_mean,   \
                                running_var, momentum, mean_shift, batches_tracked, partials, sums, st)
    if (x_dtype == AOPT_F32 && out_dtype == AOPT_F32) AOPT_BN_FWD(float, float);
    if (x_dtype == AOPT_F32 && out_dtype == AOPT_BF16) AOPT_BN_FWD(float, __nv_bfloat16);
    if (x_dtype == AOPT_BF16 && out_dtype == AOPT_F32) AOPT_BN_FWD(__nv_bfloat16, float);
    AOPT_BN_FWD(__nv_bfloat16, __nv_bfloat16);
#undef AOPT_BN_FWD
}

extern "C" int aopt_bn_act_backward(int64_t rows, int c, const void *grad_out, const void *out, int out_dtype, const void *x,
                                    int64_t ldx, int x_dtype, const float *stats, const float *gamma, const float *row_scale,
                                    void *grad_x, int64_t ldgx, void *grad_residual, float *grad_gamma, float *grad_beta,
                                    void *workspace, size_t workspace_bytes, aopt_stream_t stream) {
    if (rows <= 0 || !bn_width_ok(c) || !grad_out || !x || !stats || !gamma || !grad_x || !grad_gamma || !grad_beta)
        return AOPT_ERR_INVALID_ARGUMENT;
    if ((x_dtype | out_dtype) & ~1) return AOPT_ERR_INVALID_ARGUMENT;
    if (ldx < c || (ldx & 3) || ldgx < c || (ldgx & 3)) return AOPT_ERR_INVALID_ARGUMENT;
    float *partials;
    double *sums;
    if (!carve_dense(workspace, workspace_bytes, 2 * c, &partials, &sums)) return AOPT_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
#define AOPT_BN_BWD(XT, OT)                                                                                               \
    return bn_backward_t<XT, OT>(rows, c, grad_out, out, x, ldx, stats, gamma, row_scale, grad_x, ldgx, grad_residual, grad_gamma,  \
                                 grad_beta, partials, sums, st)
    if (x_dtype == AOPT_F32 && out_dtype == AOPT_F32) AOPT_BN_BWD(float, float);
    if (x_dtype == AOPT_F32 && out_dtype == AOPT_BF16) AOPT_BN_BWD(float, __nv_bfloat16);
    if (x_dtype == AOPT_BF16 && out_dtype == AOPT_F32) AOPT_BN_BWD(__nv_bfloat16, float);
    AOPT_BN_BWD(__nv_bfloat16, __nv_bfloat16);
#undef AOPT_BN_BWD
}

static bool tail_src_ok(int64_t rows, const float *rel, const float *kp, const float *qp, const int *idx, int nsample) {
    return rel || (kp && qp && idx && nsample > 0 && rows % nsample == 0);
}

extern "C" int aopt_we_tail_forward(int64_t rows, int g, const float *rel_in, const float *kp, const float *qp, const int *idx,
                                    int nsample, const float *upe, const float *cst,
                                    const float *gamma, const float *beta, float eps, const float *w2, const float *b2,
                                    float *logits, float *stats_out, float *running_mean, float *running_var,
                                    float momentum, long long *batches_tracked, void *workspace, size_t workspace_bytes,
                                    aopt_stream_t stream) {
    if (rows <= 0 || !tail_src_ok(rows, rel_in, kp, qp, idx, nsample) || !gamma || !beta || !w2 || !logits || !stats_out)
        return AOPT_ERR_INVALID_ARGUMENT;
    if (!aopt_we_tail_supported(g)) return AOPT_ERR_UNSUPPORTED;
    const TailSrc rel = {rel_in, kp, qp, idx, nsample};
    float *partials;
    double *sums;
    if (!carve_dense(workspace, workspace_bytes, 3 * g + g * g, &partials, &sums)) return AOPT_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    if (g == 6)
        return we_forward_t<6, 6>(rows, rel, upe, cst, gamma, beta, eps, w2, b2, logits, stats_out, running_mean, running_var,
                                  momentum, batches_tracked, partials, sums, st);
    return we_forward_t<12, 4>(rows, rel, upe, cst, gamma, beta, eps, w2, b2, logits, stats_out, running_mean, running_var,
                               momentum, batches_tracked, partials, sums, st);
}

extern "C" int aopt_we_tail_backward(int64_t rows, int g, const float *rel_in, const float *kp, const float *qp, const int *idx,
                                     int nsample, const float *upe, const float *cst,
                                     const float *grad_logits, const float *stats, const float *gamma, const float *beta,
                                     const float *w2, float *grad_u, float *grad_gamma, float *grad_beta, float *grad_b2,
                                     float *grad_w2, void *workspace, size_t workspace_bytes, aopt_stream_t stream) {
    if (rows <= 0 || !tail_src_ok(rows, rel_in, kp, qp, idx, nsample) || !grad_logits || !stats || !gamma || !beta || !w2 ||
        !grad_u || !grad_gamma || !grad_beta || !grad_b2 || !grad_w2)
        return AOPT_ERR_INVALID_ARGUMENT;
    if (!aopt_we_tail_supported(g)) return AOPT_ERR_UNSUPPORTED;
    const TailSrc rel = {rel_in, kp, qp, idx, nsample};
    float *partials;
    double *sums;
    if (!carve_dense(workspace, workspace_bytes, 3 * g + g * g, &partials, &sums)) return AOPT_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    if (g == 6)
        return we_backward_t<6, 6>(rows, rel, upe, cst, grad_logits, stats, gamma, beta, w2, grad_u, grad_gamma, grad_beta,
                                   grad_b2, grad_w2, partials, sums, st);
    return we_backward_t<12, 4>(rows, rel, upe, cst, grad_logits, stats, gamma, beta, w2, grad_u, grad_gamma, grad_beta,
                                grad_b2, grad_w2, partials, sums, st);
}

// out (c floats) = column sums of x (rows, c) [row stride ldx]: the bias gradient of a Linear layer.
extern "C" int aopt_col_sum(int64_t rows, int c, const void *x, int64_t ldx, int x_dtype, float *out, void *workspace,
                            size_t workspace_bytes, aopt_stream_t stream) {
    if (rows <= 0 || !bn_width_ok(c) || !x || !out || (x_dtype & ~1) || ldx < c || (ldx & 3)) return AOPT_ERR_INVALID_ARGUMENT;
    float *partials;
    double *sums;
    if (!carve_dense(workspace, workspace_bytes, 2 * c, &partials, &sums)) return AOPT_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    const int grid = bn_grid_v(rows, c, 4);
    const bool pdl = tuning(kTunePdl) != 2;
    const bool v8 = bn_vec(c) == 8;
    if (x_dtype == AOPT_F32) {
        if (v8) bn_partial_kernel<float, 8><<<grid, kDenseBlock, 0, st>>>(rows, c, static_cast<const float *>(x), ldx, partials);
        else bn_partial_kernel<float, 4><<<grid, kDenseBlock, 0, st>>>(rows, c, static_cast<const float *>(x), ldx, partials);
    } else {
        if (v8) bn_partial_kernel<__nv_bfloat16, 8><<<grid, kDenseBlock, 0, st>>>(rows, c, static_cast<const __nv_bfloat16 *>(x), ldx, partials);
        else bn_partial_kernel<__nv_bfloat16, 4><<<grid, kDenseBlock, 0, st>>>(rows, c, static_cast<const __nv_bfloat16 *>(x), ldx, partials);
    }
    launch_chain(pdl, partials_reduce_kernel, div_up(c, 32), 32 * kReduceSlices, 0, st, grid, c, 2 * c, (const float *)partials,
                 (double *)nullptr, out);
    return check_launch(2);
}

extern "C" int aopt_copy_cols(int64_t rows, int c, const void *src, int64_t ld_src, int src_dtype, const float *bias, void *dst,
                              int64_t ld_dst, int dst_dtype, aopt_stream_t stream) {
    if (rows <= 0 || !bn_width_ok(c) || !src || !dst || ((src_dtype | dst_dtype) & ~1) || ld_src < c || (ld_src & 3) ||
        ld_dst < c || (ld_dst & 3))
        return AOPT_ERR_INVALID_ARGUMENT;
    cudaStream_t st = as_stream(stream);
    const int grid = bn_grid(rows, c);
#define AOPT_COPY(ST, DT)                                                                                          \
    copy_cols_kernel<ST, DT><<<grid, kDenseBlock, 0, st>>>(rows, c, static_cast<const ST *>(src), ld_src, bias,  \
                                                           static_cast<DT *>(dst), ld_dst)
    if (src_dtype == AOPT_F32 && dst_dtype == AOPT_F32) AOPT_COPY(float, float);
    else if (src_dtype == AOPT_F32) AOPT_COPY(float, __nv_bfloat16);
    else if (dst_dtype == AOPT_F32) AOPT_COPY(__nv_bfloat16, float);
    else AOPT_COPY(__nv_bfloat16, __nv_bfloat16);
#undef AOPT_COPY
    return check_launch(1);
}

// widths served: 3 (relative coordinates into linear_p_bias[0]), the in_channels / num_classes / groups of the configs
// BASELINE.json names (S3DIS 6 / 13, ScanNet 9 / 20, SemanticKITTI 4 / 19; weight_encoding groups 6 / 12)
#define AOPT_SKINNY_WIDTHS(X) X(3) X(4) X(6) X(9) X(12) X(13) X(19) X(20)
extern "C" int aopt_skinny_wgrad_supported(int g, int c) {
    if (!bn_width_ok(c)) return 0;
#define AOPT_CASE(GG) if (g == GG) return 1;
    AOPT_SKINNY_WIDTHS(AOPT_CASE)
#undef AOPT_CASE
    return 0;
}

// out (g, c) fp32 = gradᵀ · x;  grad (rows, g) dense, x (rows, c) with row stride ldx.  workspace: aopt_dense_workspace_bytes(g * c).
extern "C" int aopt_skinny_wgrad(int64_t rows, int g, int c, const void *grad, int grad_dtype, const void *x, int64_t ldx,
                                 int x_dtype, float *out, void *workspace, size_t workspace_bytes, aopt_stream_t stream) {
    if (rows <= 0 || !grad || !x || !out || ((grad_dtype | x_dtype) & ~1) || ldx < c || (ldx & 3)) return AOPT_ERR_INVALID_ARGUMENT;
    if (!aopt_skinny_wgrad_supported(g, c)) return AOPT_ERR_UNSUPPORTED;
    float *partials;
    double *sums;
    if (!carve_dense(workspace, workspace_bytes, g * c, &partials, &sums)) return AOPT_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    const int grid = bn_grid(rows, c);
    const bool pdl = tuning(kTunePdl) != 2;
#define AOPT_SW(GT, XT, GG)                                                                                         \
    skinny_wgrad_kernel<GT, XT, GG><<<grid, kDenseBlock, 0, st>>>(rows, c, static_cast<const GT *>(grad),        \
                                                                  static_cast<const XT *>(x), ldx, partials)
#define AOPT_CASE(GG)                                                                   \
    if (g == GG) {                                                                      \
        if (grad_dtype == AOPT_F32 && x_dtype == AOPT_F32) AOPT_SW(float, float, GG);   \
        else if (grad_dtype == AOPT_F32) AOPT_SW(float, __nv_bfloat16, GG);             \
        else if (x_dtype == AOPT_F32) AOPT_SW(__nv_bfloat16, float, GG);                \
        else AOPT_SW(__nv_bfloat16, __nv_bfloat16, GG);                                 \
    }
    AOPT_SKINNY_WIDTHS(AOPT_CASE)
#undef AOPT_CASE
#undef AOPT_SW
    launch_chain(pdl, partials_reduce_kernel, div_up(g * c, 32), 32 * kReduceSlices, 0, st, grid, g * c, g * c,
                 (const float *)partials, (double *)nullptr, out);
    return check_launch(2);
}

// out (rows, g) fp32 = x (rows, c) · wᵀ (+ bias), w (g, c) fp32 (aopt_skinny_wgrad_supported).
extern "C" int aopt_skinny_linear(int64_t rows, int g, int c, const void *x, int64_t ldx, int x_dtype, const float *w,
                                  const float *bias, float *out, aopt_stream_t stream) {
    if (rows <= 0 || !x || !w || !out || (x_dtype & ~1) || ldx < c || (ldx & 3)) return AOPT_ERR_INVALID_ARGUMENT;
    if (!aopt_skinny_wgrad_supported(g, c)) return AOPT_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    const int grid = row_grid(rows);
    const size_t smem = (size_t)g * c * sizeof(float);
    if (smem > 48 * 1024) return AOPT_ERR_UNSUPPORTED;
#define AOPT_CASE(GG)                                                                                                            \
    if (g == GG) {                                                                                                               \
        if (x_dtype == AOPT_F32)                                                                                                 \
            skinny_linear_kernel<float, GG><<<grid, kDenseBlock, smem, st>>>(rows, c, static_cast<const float *>(x), ldx, w, bias, out); \
        else                                                                                                                     \
            skinny_linear_kernel<__nv_bfloat16, GG><<<grid, kDenseBlock, smem, st>>>(rows, c, static_cast<const __nv_bfloat16 *>(x), ldx, w, bias, out); \
    }
    AOPT_SKINNY_WIDTHS(AOPT_CASE)
#undef AOPT_CASE
    return check_launch(1);
}

// grad_x (rows, c) [row stride ldgx, fp32 or bf16] = grad (rows, g) fp32 · w (g, c).
extern "C" int aopt_skinny_dgrad(int64_t rows, int g, int c, const float *grad, const float *w, void *grad_x, int64_t ldgx,
                                 int x_dtype, aopt_stream_t stream) {
    if (rows <= 0 || !grad || !w || !grad_x || (x_dtype & ~1) || ldgx < c || (ldgx & 3)) return AOPT_ERR_INVALID_ARGUMENT;
    if (!aopt_skinny_wgrad_supported(g, c)) return AOPT_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    const int grid = bn_grid(rows, c);
#define AOPT_CASE(GG)                                                                                                                  \
    if (g == GG) {                                                                                                                     \
        if (x_dtype == AOPT_F32) skinny_dgrad_kernel<float, GG><<<grid, kDenseBlock, 0, st>>>(rows, c, grad, w, static_cast<float *>(grad_x), ldgx); \
        else skinny_dgrad_kernel<__nv_bfloat16, GG><<<grid, kDenseBlock, 0, st>>>(rows, c, grad, w, static_cast<__nv_bfloat16 *>(grad_x), ldgx);      \
    }
    AOPT_SKINNY_WIDTHS(AOPT_CASE)
#undef AOPT_CASE
    return check_launch(1);
}
