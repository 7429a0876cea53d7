// grouping.cu — neighbour gather and its atomic-free segmented scatter-add backward.
//
// Replaces (paths relative to /root/reference):
//   libs/pointops/src/grouping/grouping_cuda_kernel.cu:5-25  (1 thread / element, 3 div-mods per
//       element, 4-byte accesses, float atomicAdd backward)
//   libs/pointops/functions/grouping.py:36-60                (cat zero row + index + sub + einsum + cat)
//   pointcept/models/point_transformer_v2/point_transformer_v2m2_base.py:109,112  (key[idx] - q)
//
// HBM layout: features are (rows, c) row-major fp32; one thread moves one 128-bit chunk of one
// (query, neighbour) row, so a warp covers 512 contiguous output bytes.  The gathered source rows
// go through the read-only path with default caching (each source row is re-read ~k times by
// nearby queries and stays in L1/L2); the streamed (m,k,c) side uses L1::no_allocate.
//
// Algorithmic bytes: forward 4·M·k + 4·N·C + 4·M·k·C;
// backward 4·M·k·C + 4·M·k + 4·(N+1) + 4·N·C.
#include <stdlib.h>

#include "common.cuh"

namespace aopt {

constexpr int kBlock = 256;

// out[p, :] = in[idx[p], :]  (zeros for idx < 0)
template <int VEC>
__global__ void __launch_bounds__(kBlock)
gather_rows_kernel(long long rows, int chunks, int c, const float *__restrict__ in,
                   const int *__restrict__ idx, float *__restrict__ out, int out_stride) {
    const long long total = rows * chunks;
    const long long step = (long long)gridDim.x * kBlock;
    for (long long t = (long long)blockIdx.x * kBlock + threadIdx.x; t < total; t += step) {
        RowCol rc = split(t, chunks);
        int j = __ldg(idx + rc.row);
        Chunk<VEC> v = Chunk<VEC>::zero();
        if (j >= 0) v = Chunk<VEC>::gather(in + (size_t)j * c + rc.col * VEC);
        v.store_stream(out + (size_t)rc.row * out_stride + rc.col * VEC);
    }
}

// out[m, s, :] = key[idx[m,s], :] - query[m, :]
template <int VEC>
__global__ void __launch_bounds__(kBlock)
gather_sub_kernel(long long rows, int nsample, int chunks, int c, const float *__restrict__ key,
                  const float *__restrict__ query, const int *__restrict__ idx,
                  float *__restrict__ out) {
    const long long total = rows * chunks;
    const long long step = (long long)gridDim.x * kBlock;
    for (long long t = (long long)blockIdx.x * kBlock + threadIdx.x; t < total; t += step) {
        RowCol rc = split(t, chunks);
        int j = __ldg(idx + rc.row);
        long long m = rc.row / nsample;
        Chunk<VEC> v = Chunk<VEC>::zero();
        if (j >= 0) v = Chunk<VEC>::gather(key + (size_t)j * c + rc.col * VEC);
        v.sub(Chunk<VEC>::gather(query + (size_t)m * c + rc.col * VEC));
        v.store_stream(out + (size_t)rc.row * c + rc.col * VEC);
    }
}

// 128-bit path of out[m, s, :] = key[idx[m,s], :] - query[m, :]: one thread per (query, 4-channel chunk)
// walking the k neighbour slots four at a time — four independent gathers and four 128-bit streaming
// stores in flight per thread (the one-item-per-thread form above is latency-bound at ~60 % of HBM
// peak).  The k rows of one query are contiguous in `out`, so a warp's stores fill whole 128-byte lines.
__global__ void __launch_bounds__(kBlock)
gather_sub_rows_kernel(long long m, int k, int chunks, int c, const float *__restrict__ key,
                       const float *__restrict__ query, const int *__restrict__ idx,
                       float *__restrict__ out) {
    const long long total = m * chunks;
    const long long step = (long long)gridDim.x * kBlock;
    for (long long t = (long long)blockIdx.x * kBlock + threadIdx.x; t < total; t += step) {
        const long long row = t / chunks;
        const int col = (int)(t - row * chunks);
        const float4 q = ldg_gather4(query + (size_t)row * c + col * 4);
        const int *ix = idx + (size_t)row * k;
        const float *kbase = key + col * 4;
        float *o = out + (size_t)row * k * c + col * 4;
        int s = 0;
        for (; s + 4 <= k; s += 4) {
            int j[4];
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) j[u] = __ldg(ix + s + u);
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ldg_gather4(kbase + (size_t)max(j[u], 0) * c);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool keep = j[u] >= 0;  // idx == -1 selects the zero row (grouping.py:41-42)
                stg_stream4(o + (size_t)(s + u) * c,
                            make_float4((keep ? v[u].x : 0.f) - q.x, (keep ? v[u].y : 0.f) - q.y,
                                        (keep ? v[u].z : 0.f) - q.z, (keep ? v[u].w : 0.f) - q.w));
            }
        }
        for (; s < k; ++s) {
            const int j = __ldg(ix + s);
            const float4 v = ldg_gather4(kbase + (size_t)max(j, 0) * c);
            const bool keep = j >= 0;
            stg_stream4(o + (size_t)s * c, make_float4((keep ? v.x : 0.f) - q.x, (keep ? v.y : 0.f) - q.y,
                                                       (keep ? v.z : 0.f) - q.z, (keep ? v.w : 0.f) - q.w));
        }
    }
}

// Compile-time neighbour count (NS in {8,16,32}): the idx row of the next item is prefetched and all NS
// gathered pieces are requested at once with cp.async into the thread's own shared-memory slots, so the
// DRAM/L2 latency is exposed once per item (NS·16 bytes of shared memory per thread).
constexpr int kSubNsBlock = 128;

template <int NS, bool CG>
__global__ void __launch_bounds__(kSubNsBlock)
gather_sub_ns_kernel(long long m, int chunks, int c, const float *__restrict__ key,
                     const float *__restrict__ query, const int *__restrict__ idx,
                     float *__restrict__ out) {
    extern __shared__ float4 stage[];  // [NS][kSubNsBlock]
    float4 *sv = stage + threadIdx.x;
    const long long total = m * chunks;
    const long long step = (long long)gridDim.x * kSubNsBlock;
    long long t = (long long)blockIdx.x * kSubNsBlock + threadIdx.x;
    int jn[NS];
    auto load_row = [&](long long row) {
        const int4 *r4 = reinterpret_cast<const int4 *>(idx + (size_t)row * NS);
#pragma unroll
        for (int u = 0; u < NS / 4; ++u) {
            const int4 q4 = __ldg(r4 + u);
            jn[4 * u] = q4.x; jn[4 * u + 1] = q4.y; jn[4 * u + 2] = q4.z; jn[4 * u + 3] = q4.w;
        }
    };
    if (t < total) load_row(t / chunks);
    for (; t < total; t += step) {
        const long long row = t / chunks;
        const int col = (int)(t - row * chunks);
        int j[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) j[s] = jn[s];
        const float *kbase = key + col * 4;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            if (CG) cp_async16_stream(sv + s * kSubNsBlock, kbase + (size_t)max(j[s], 0) * c);
            else cp_async16_gather(sv + s * kSubNsBlock, kbase + (size_t)max(j[s], 0) * c);
        }
        cp_async_commit();
        const float4 q = ldg_gather4(query + (size_t)row * c + col * 4);
        if (t + step < total) load_row((t + step) / chunks);
        float *o = out + (size_t)row * NS * c + col * 4;
        cp_async_wait_all();
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const float4 v = sv[s * kSubNsBlock];
            const bool keep = j[s] >= 0;  // idx == -1 selects the zero row (grouping.py:41-42)
            stg_stream4(o + (size_t)s * c, make_float4((keep ? v.x : 0.f) - q.x, (keep ? v.y : 0.f) - q.y,
                                                       (keep ? v.z : 0.f) - q.z, (keep ? v.w : 0.f) - q.w));
        }
    }
}

template <int NS>
static void launch_gather_sub_ns(long long m, int chunks, int c, const float *key, const float *query,
                                 const int *idx, float *out, cudaStream_t st) {
    const size_t smem = (size_t)NS * 16 * kSubNsBlock;
    static bool once = (cudaFuncSetAttribute(gather_sub_ns_kernel<NS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        cudaFuncSetAttribute(gather_sub_ns_kernel<NS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), true);
    (void)once;
    // Measured at level 0 (NS = 16): L2-only copies (.cg) 204 us at 3 CTAs/SM, 235 at 4, 247 at 8; L1-allocating
    // copies (.ca) 298-329 us (the 61 MB of key rows thrash L1 in S3DIS point order); rows kernel 253 us.
    int per_sm = NS <= 8 ? 6 : NS <= 16 ? 3 : 2;
    static const int env_ctas = [] { const char *e = getenv("AOPT_GATHER_SUB_CTAS"); return e ? atoi(e) : 0; }();
    static const bool cg = [] { const char *e = getenv("AOPT_GATHER_SUB_CG"); return !(e && e[0] == '0'); }();
    if (env_ctas >= 1 && env_ctas <= 16) per_sm = env_ctas;
    if (cg)
        gather_sub_ns_kernel<NS, true><<<stride_grid(m * chunks, kSubNsBlock, per_sm), kSubNsBlock, smem, st>>>(m, chunks, c, key, query, idx, out);
    else
        gather_sub_ns_kernel<NS, false><<<stride_grid(m * chunks, kSubNsBlock, per_sm), kSubNsBlock, smem, st>>>(m, chunks, c, key, query, idx, out);
}

// grad_in[j, :] = scale * sum_{e in row j} grad_out[perm[e], :]   — one thread per (source row,
// chunk); entries are visited in ascending flat position, so the sum order is fixed.
template <int VEC>
__global__ void __launch_bounds__(kBlock)
segmented_sum_kernel(long long n, int chunks, int c, const float *__restrict__ grad_out,
                     int go_stride, const int *__restrict__ rowptr, const int *__restrict__ perm,
                     float scale, float *__restrict__ grad_in) {
    const long long total = n * chunks;
    const long long step = (long long)gridDim.x * kBlock;
    for (long long t = (long long)blockIdx.x * kBlock + threadIdx.x; t < total; t += step) {
        RowCol rc = split(t, chunks);
        int e = __ldg(rowptr + rc.row);
        const int e_end = __ldg(rowptr + rc.row + 1);
        const float *base = grad_out + rc.col * VEC;
        Chunk<VEC> acc = Chunk<VEC>::zero();
        // 4 independent 128-bit loads in flight per thread
        for (; e + 4 <= e_end; e += 4) {
            int p0 = __ldg(perm + e), p1 = __ldg(perm + e + 1), p2 = __ldg(perm + e + 2),
                p3 = __ldg(perm + e + 3);
            Chunk<VEC> a0 = Chunk<VEC>::stream(base + (size_t)p0 * go_stride);
            Chunk<VEC> a1 = Chunk<VEC>::stream(base + (size_t)p1 * go_stride);
            Chunk<VEC> a2 = Chunk<VEC>::stream(base + (size_t)p2 * go_stride);
            Chunk<VEC> a3 = Chunk<VEC>::stream(base + (size_t)p3 * go_stride);
            acc.add(a0); acc.add(a1); acc.add(a2); acc.add(a3);
        }
        for (; e < e_end; ++e) {
            int p = __ldg(perm + e);
            acc.add(Chunk<VEC>::stream(base + (size_t)p * go_stride));
        }
        acc.scale(scale);
        acc.store(grad_in + (size_t)rc.row * c + rc.col * VEC);
    }
}

// out[m, :] = scale * sum_s grad[m, s, :]
template <int VEC>
__global__ void __launch_bounds__(kBlock)
sum_over_k_kernel(long long m, int nsample, int chunks, int c, const float *__restrict__ grad,
                  float scale, float *__restrict__ out) {
    const long long total = m * chunks;
    const long long step = (long long)gridDim.x * kBlock;
    for (long long t = (long long)blockIdx.x * kBlock + threadIdx.x; t < total; t += step) {
        RowCol rc = split(t, chunks);
        const float *base = grad + (size_t)rc.row * nsample * c + rc.col * VEC;
        Chunk<VEC> acc = Chunk<VEC>::zero();
        int s = 0;
        for (; s + 4 <= nsample; s += 4) {
            Chunk<VEC> a0 = Chunk<VEC>::stream(base + (size_t)(s + 0) * c);
            Chunk<VEC> a1 = Chunk<VEC>::stream(base + (size_t)(s + 1) * c);
            Chunk<VEC> a2 = Chunk<VEC>::stream(base + (size_t)(s + 2) * c);
            Chunk<VEC> a3 = Chunk<VEC>::stream(base + (size_t)(s + 3) * c);
            acc.add(a0); acc.add(a1); acc.add(a2); acc.add(a3);
        }
        for (; s < nsample; ++s) acc.add(Chunk<VEC>::stream(base + (size_t)s * c));
        acc.scale(scale);
        acc.store(out + (size_t)rc.row * c + rc.col * VEC);
    }
}

// Backward of relation = key[idx] - query[:, None] over ONE point set (queries == sources, the GVA case):
//   grad_query[j] = -sum_s grad[j, s, :]          (the k rows of query j: 16 x C contiguous floats)
//   grad_key[j]   =  sum_{e in row j} grad[perm[e]] (the rows that gathered j)
// segmented_sum_kernel + sum_over_k_kernel read the (N,k,C) gradient twice from HBM (2 x 983 MB at level 0).
// Fused, every row is still read twice, but the two reads fall close together in time: neighbours are close
// in index (median |i - j| = 559, 90 % < 1643 in S3DIS order, profiles/r01h_bv_locality.md) and a grid-stride
// loop moves the whole chip through one compact window of rows, so the second read of a row hits L2.
// The grid is kept small (ctas_per_sm x 148 CTAs) to keep that window a fraction of the 126 MB L2.
template <int VEC>
__global__ void __launch_bounds__(kBlock)
relation_backward_kernel(long long n, int nsample, int chunks, int c, const float *__restrict__ grad,
                         const int *__restrict__ rowptr, const int *__restrict__ perm,
                         float *__restrict__ grad_key, float *__restrict__ grad_query) {
    const long long total = n * chunks;
    const long long step = (long long)gridDim.x * kBlock;
    for (long long t = (long long)blockIdx.x * kBlock + threadIdx.x; t < total; t += step) {
        RowCol rc = split(t, chunks);
        int e = __ldg(rowptr + rc.row);
        const int e_end = __ldg(rowptr + rc.row + 1);
        // grad_query: the query's own k rows (streamed; brings them into L2 for the row walks nearby)
        const float *own = grad + (size_t)rc.row * nsample * c + rc.col * VEC;
        Chunk<VEC> accq = Chunk<VEC>::zero();
        int s = 0;
        for (; s + 4 <= nsample; s += 4) {
            Chunk<VEC> a0 = Chunk<VEC>::stream(own + (size_t)(s + 0) * c);
            Chunk<VEC> a1 = Chunk<VEC>::stream(own + (size_t)(s + 1) * c);
            Chunk<VEC> a2 = Chunk<VEC>::stream(own + (size_t)(s + 2) * c);
            Chunk<VEC> a3 = Chunk<VEC>::stream(own + (size_t)(s + 3) * c);
            accq.add(a0); accq.add(a1); accq.add(a2); accq.add(a3);
        }
        for (; s < nsample; ++s) accq.add(Chunk<VEC>::stream(own + (size_t)s * c));
        accq.scale(-1.0f);
        accq.store(grad_query + (size_t)rc.row * c + rc.col * VEC);
        // grad_key: in-edges in ascending flat position (same order as segmented_sum_kernel)
        const float *base = grad + rc.col * VEC;
        Chunk<VEC> acc = Chunk<VEC>::zero();
        for (; e + 4 <= e_end; e += 4) {
            int p0 = __ldg(perm + e), p1 = __ldg(perm + e + 1), p2 = __ldg(perm + e + 2), p3 = __ldg(perm + e + 3);
            Chunk<VEC> a0 = Chunk<VEC>::stream(base + (size_t)p0 * c);
            Chunk<VEC> a1 = Chunk<VEC>::stream(base + (size_t)p1 * c);
            Chunk<VEC> a2 = Chunk<VEC>::stream(base + (size_t)p2 * c);
            Chunk<VEC> a3 = Chunk<VEC>::stream(base + (size_t)p3 * c);
            acc.add(a0); acc.add(a1); acc.add(a2); acc.add(a3);
        }
        for (; e < e_end; ++e) acc.add(Chunk<VEC>::stream(base + (size_t)__ldg(perm + e) * c));
        acc.store(grad_key + (size_t)rc.row * c + rc.col * VEC);
    }
}

// 128-bit path: per iteration a thread requests FOUR of its own rows and FOUR gathered rows at once with
// cp.async (L2-only) into its own shared-memory slots — eight 16-byte requests in flight per thread without
// holding registers (ptxas folds register-destination batches of this shape back into ~3 live loads) — and
// prefetches the next four perm entries while they fly.  Slots past the end of either list re-request one of the
// thread's own rows (an L2 hit) and are dropped by a select, so the summation orders are exactly those of
// segmented_sum_kernel / sum_over_k_kernel.
constexpr int kRelBlock = 256;

__global__ void __launch_bounds__(kRelBlock)
relation_backward_vec_kernel(long long n, int nsample, int chunks, int c, const float *__restrict__ grad,
                             const int *__restrict__ rowptr, const int *__restrict__ perm,
                             float *__restrict__ grad_key, float *__restrict__ grad_query) {
    __shared__ float4 stage[8 * kRelBlock];
    float4 *sg = stage + threadIdx.x;
    const long long total = n * chunks;
    const long long step = (long long)gridDim.x * kRelBlock;
    for (long long t = (long long)blockIdx.x * kRelBlock + threadIdx.x; t < total; t += step) {
        RowCol rc = split(t, chunks);
        const int e0 = __ldg(rowptr + rc.row), e_end = __ldg(rowptr + rc.row + 1);
        const long long own0 = rc.row * nsample;          // flat row of (query, slot 0)
        const float *base = grad + rc.col * 4;
        const int iters = max((nsample + 3) >> 2, (e_end - e0 + 3) >> 2);
        long long pn[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) pn[u] = (e0 + u < e_end) ? (long long)__ldg(perm + e0 + u) : own0;
        float4 accq = make_float4(0.f, 0.f, 0.f, 0.f), acck = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int it = 0; it < iters; ++it) {
            const int s = it * 4, e = e0 + it * 4;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                cp_async16_stream(sg + u * kRelBlock, base + (size_t)(own0 + min(s + u, nsample - 1)) * c);
#pragma unroll
            for (int u = 0; u < 4; ++u) cp_async16_stream(sg + (4 + u) * kRelBlock, base + (size_t)pn[u] * c);
            cp_async_commit();
#pragma unroll
            for (int u = 0; u < 4; ++u) pn[u] = (e + 4 + u < e_end) ? (long long)__ldg(perm + e + 4 + u) : own0;
            cp_async_wait_all();
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 a = sg[u * kRelBlock];
                const bool keep = s + u < nsample;
                accq.x = keep ? accq.x + a.x : accq.x; accq.y = keep ? accq.y + a.y : accq.y;
                accq.z = keep ? accq.z + a.z : accq.z; accq.w = keep ? accq.w + a.w : accq.w;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 g4 = sg[(4 + u) * kRelBlock];
                const bool keep = e + u < e_end;
                acck.x = keep ? acck.x + g4.x : acck.x; acck.y = keep ? acck.y + g4.y : acck.y;
                acck.z = keep ? acck.z + g4.z : acck.z; acck.w = keep ? acck.w + g4.w : acck.w;
            }
        }
        accq.x *= -1.0f; accq.y *= -1.0f; accq.z *= -1.0f; accq.w *= -1.0f;
        *reinterpret_cast<float4 *>(grad_query + (size_t)rc.row * c + rc.col * 4) = accq;
        *reinterpret_cast<float4 *>(grad_key + (size_t)rc.row * c + rc.col * 4) = acck;
    }
}

// out[p, 0..2] = (xyz[idx[p]] - new_xyz[p / nsample]) * sign(idx[p] + 1)
__global__ void __launch_bounds__(kBlock)
group_xyz_kernel(long long rows, int nsample, const float *__restrict__ xyz,
                 const float *__restrict__ new_xyz, const int *__restrict__ idx,
                 float *__restrict__ out, int out_stride) {
    const long long step = (long long)gridDim.x * kBlock;
    for (long long p = (long long)blockIdx.x * kBlock + threadIdx.x; p < rows; p += step) {
        int j = __ldg(idx + p);
        long long m = p / nsample;
        float rx = 0.f, ry = 0.f, rz = 0.f;
        float qx = __ldg(new_xyz + m * 3 + 0), qy = __ldg(new_xyz + m * 3 + 1),
              qz = __ldg(new_xyz + m * 3 + 2);
        if (j >= 0) {
            rx = __ldg(xyz + (size_t)j * 3 + 0) - qx;
            ry = __ldg(xyz + (size_t)j * 3 + 1) - qy;
            rz = __ldg(xyz + (size_t)j * 3 + 2) - qz;
        } else {
            // grouping.py:41-57: the padded zero row gives (0 - q) * 0 = -0.0 or +0.0 by sign of q
            rx = (0.f - qx) * 0.f; ry = (0.f - qy) * 0.f; rz = (0.f - qz) * 0.f;
        }
        float *o = out + (size_t)p * out_stride;
        o[0] = rx; o[1] = ry; o[2] = rz;
    }
}

// Packed (m, k, 3) output with k a multiple of 4: one thread per (query, four slots) — one 128-bit idx
// load, twelve gathered coordinates, three 128-bit stores (the per-slot kernel above writes 12-byte
// pieces with scalar stores, three store instructions over the same lines).
__global__ void __launch_bounds__(kBlock)
group_xyz_packed_kernel(long long m, int k4, const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                        const int *__restrict__ idx, float *__restrict__ out) {
    const long long total = m * k4;
    const long long step = (long long)gridDim.x * kBlock;
    for (long long t = (long long)blockIdx.x * kBlock + threadIdx.x; t < total; t += step) {
        const long long row = t / k4;
        const int4 j4 = __ldg(reinterpret_cast<const int4 *>(idx) + t);
        const int j[4] = {j4.x, j4.y, j4.z, j4.w};
        const float qx = __ldg(new_xyz + row * 3 + 0), qy = __ldg(new_xyz + row * 3 + 1),
                    qz = __ldg(new_xyz + row * 3 + 2);
        float r[12];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float *p = xyz + (size_t)max(j[u], 0) * 3;
            const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
            // grouping.py:41-57: the padded zero row gives (0 - q) * 0 = -0.0 or +0.0 by the sign of q
            r[3 * u + 0] = j[u] >= 0 ? x - qx : (0.f - qx) * 0.f;
            r[3 * u + 1] = j[u] >= 0 ? y - qy : (0.f - qy) * 0.f;
            r[3 * u + 2] = j[u] >= 0 ? z - qz : (0.f - qz) * 0.f;
        }
        float *o = out + (size_t)t * 12;
        stg_stream4(o, make_float4(r[0], r[1], r[2], r[3]));
        stg_stream4(o + 4, make_float4(r[4], r[5], r[6], r[7]));
        stg_stream4(o + 8, make_float4(r[8], r[9], r[10], r[11]));
    }
}

}  // namespace aopt

using namespace aopt;

extern "C" int aopt_grouping_forward(int m, int nsample, int c, const float *input, const int *idx,
                                     float *output, int out_stride, aopt_stream_t stream) {
    if (m < 0 || nsample < 1 || c < 1 || out_stride < c) return AOPT_ERR_INVALID_ARGUMENT;
    long long rows = (long long)m * nsample;
    if (rows == 0) return AOPT_OK;
    if (!input || !idx || !output) return AOPT_ERR_INVALID_ARGUMENT;
    bool vec = (c % 4 == 0) && (out_stride % 4 == 0) && aligned16(input) && aligned16(output);
    if (vec) {
        int chunks = c / 4;
        gather_rows_kernel<4><<<stride_grid(rows * chunks, kBlock, 8), kBlock, 0, as_stream(stream)>>>(
            rows, chunks, c, input, idx, output, out_stride);
    } else {
        gather_rows_kernel<1><<<stride_grid(rows * c, kBlock, 8), kBlock, 0, as_stream(stream)>>>(
            rows, c, c, input, idx, output, out_stride);
    }
    return check_launch();
}

extern "C" int aopt_grouping_backward(int n, int c, const float *grad_output, int go_stride,
                                      const int *rowptr, const int *perm, float scale,
                                      float *grad_input, aopt_stream_t stream) {
    if (n < 0 || c < 1 || go_stride < c) return AOPT_ERR_INVALID_ARGUMENT;
    if (n == 0) return AOPT_OK;
    if (!grad_output || !rowptr || !perm || !grad_input) return AOPT_ERR_INVALID_ARGUMENT;
    bool vec = (c % 4 == 0) && (go_stride % 4 == 0) && aligned16(grad_output) && aligned16(grad_input);
    // (The batched csr_walk.cuh kernel is used by the WEIGHTED walks only: for this plain streaming sum it
    // measured 270-384 us against 212 us at level 0 — its end-of-row filler loads all stream the same
    // row 0 from L2 and ptxas folds the eight loads back into three registers.)
    if (vec) {
        int chunks = c / 4;
        segmented_sum_kernel<4><<<stride_grid((long long)n * chunks, kBlock, 8), kBlock, 0, as_stream(stream)>>>(
            n, chunks, c, grad_output, go_stride, rowptr, perm, scale, grad_input);
    } else {
        segmented_sum_kernel<1><<<stride_grid((long long)n * c, kBlock, 8), kBlock, 0, as_stream(stream)>>>(
            n, c, c, grad_output, go_stride, rowptr, perm, scale, grad_input);
    }
    return check_launch();
}

// grad (n,nsample,c) of relation = key[idx] - query[:,None] with queries == sources (n rows each):
// grad_key (n,c) through the CSR of idx, grad_query (n,c) = -sum over the slots.  One pass (see kernel).
extern "C" int aopt_relation_backward(int n, int nsample, int c, const float *grad, const int *rowptr,
                                      const int *perm, float *grad_key, float *grad_query,
                                      aopt_stream_t stream) {
    if (n < 0 || nsample < 1 || c < 1) return AOPT_ERR_INVALID_ARGUMENT;
    if (n == 0) return AOPT_OK;
    if (!grad || !rowptr || !perm || !grad_key || !grad_query) return AOPT_ERR_INVALID_ARGUMENT;
    static const int ctas = [] { const char *e = getenv("AOPT_RELBWD_CTAS"); int v = e ? atoi(e) : 0; return v >= 1 && v <= 8 ? v : 4; }();
    const bool vec = (c % 4 == 0) && aligned16(grad) && aligned16(grad_key) && aligned16(grad_query);
    if (vec) {
        const int chunks = c / 4;
        relation_backward_vec_kernel<<<stride_grid((long long)n * chunks, kRelBlock, ctas), kRelBlock, 0, as_stream(stream)>>>(
            n, nsample, chunks, c, grad, rowptr, perm, grad_key, grad_query);
    } else {
        relation_backward_kernel<1><<<stride_grid((long long)n * c, kBlock, ctas), kBlock, 0, as_stream(stream)>>>(
            n, nsample, c, c, grad, rowptr, perm, grad_key, grad_query);
    }
    return check_launch();
}

extern "C" int aopt_group_xyz(int m, int nsample, const float *xyz, const float *new_xyz,
                              const int *idx, float *out, int out_stride, aopt_stream_t stream) {
    if (m < 0 || nsample < 1 || out_stride < 3) return AOPT_ERR_INVALID_ARGUMENT;
    long long rows = (long long)m * nsample;
    if (rows == 0) return AOPT_OK;
    if (!xyz || !new_xyz || !idx || !out) return AOPT_ERR_INVALID_ARGUMENT;
    if (out_stride == 3 && nsample % 4 == 0 && aligned16(idx) && aligned16(out)) {
        const int k4 = nsample / 4;
        group_xyz_packed_kernel<<<stride_grid((long long)m * k4, kBlock, 8), kBlock, 0, as_stream(stream)>>>(
            m, k4, xyz, new_xyz, idx, out);
    } else {
        group_xyz_kernel<<<stride_grid(rows, kBlock, 8), kBlock, 0, as_stream(stream)>>>(
            rows, nsample, xyz, new_xyz, idx, out, out_stride);
    }
    return check_launch();
}

extern "C" int aopt_gather_sub_forward(int m, int nsample, int c, const float *key,
                                       const float *query, const int *idx, float *out,
                                       aopt_stream_t stream) {
    if (m < 0 || nsample < 1 || c < 1) return AOPT_ERR_INVALID_ARGUMENT;
    long long rows = (long long)m * nsample;
    if (rows == 0) return AOPT_OK;
    if (!key || !query || !idx || !out) return AOPT_ERR_INVALID_ARGUMENT;
    bool vec = (c % 4 == 0) && aligned16(key) && aligned16(query) && aligned16(out);
    if (vec) {
        int chunks = c / 4;
        // Default = the cp.async NS kernel with L2-only copies (see launch_gather_sub_ns for the measurements);
        // AOPT_GATHER_SUB_IMPL=rows selects the register kernel (also used for other neighbour counts).
        static const bool use_ns = [] { const char *e = getenv("AOPT_GATHER_SUB_IMPL"); return !(e && e[0] == 'r'); }();
        const bool ns_ok = aligned16(idx) && use_ns;
        if (ns_ok && nsample == 16) launch_gather_sub_ns<16>(m, chunks, c, key, query, idx, out, as_stream(stream));
        else if (ns_ok && nsample == 8) launch_gather_sub_ns<8>(m, chunks, c, key, query, idx, out, as_stream(stream));
        else if (ns_ok && nsample == 32) launch_gather_sub_ns<32>(m, chunks, c, key, query, idx, out, as_stream(stream));
        else
            gather_sub_rows_kernel<<<stride_grid((long long)m * chunks, kBlock, 8), kBlock, 0, as_stream(stream)>>>(
                m, nsample, chunks, c, key, query, idx, out);
    } else {
        gather_sub_kernel<1><<<stride_grid(rows * c, kBlock, 8), kBlock, 0, as_stream(stream)>>>(
            rows, nsample, c, c, key, query, idx, out);
    }
    return check_launch();
}

extern "C" int aopt_sum_over_k(int m, int nsample, int c, const float *grad, float scale,
                               float *out, aopt_stream_t stream) {
    if (m < 0 || nsample < 1 || c < 1) return AOPT_ERR_INVALID_ARGUMENT;
    if (m == 0) return AOPT_OK;
    if (!grad || !out) return AOPT_ERR_INVALID_ARGUMENT;
    bool vec = (c % 4 == 0) && aligned16(grad) && aligned16(out);
    if (vec) {
        int chunks = c / 4;
        sum_over_k_kernel<4><<<stride_grid((long long)m * chunks, kBlock, 8), kBlock, 0, as_stream(stream)>>>(
            m, nsample, chunks, c, grad, scale, out);
    } else {
        sum_over_k_kernel<1><<<stride_grid((long long)m * c, kBlock, 8), kBlock, 0, as_stream(stream)>>>(
            m, nsample, c, c, grad, scale, out);
    }
    return check_launch();
}
