// interp.cu — three-NN inverse-distance interpolation, forward and backward.
//
// Replaces /root/reference/libs/pointops/functions/interpolation.py:8-22 (k separate
// gather+mul+add passes over (n,c), backward = k index_put_(accumulate)) and the CUDA variant
// libs/pointops/src/interpolation/interpolation_cuda_kernel.cu:5-33 (1 thread per (n,c), float
// atomicAdd backward).  The neighbour search itself is aopt_knn_query with nsample = k.
//
// Bytes: forward 24·Nf + 4·Nc·C + 4·Nf·C; backward 4·Nf·C + 36·Nf + 4(Nc+1) + 4·Nc·C.
#include "common.cuh"
#include "csr_walk.cuh"

namespace aopt {

constexpr int kInterpBlock = 256;

// weight[n,i] = r_i / sum_j r_j with r = 1/(sqrt(d2)+1e-8): every operation IEEE-rounded fp32
// (sqrt, add, reciprocal, sequential sum, divide) like the torch ops of interpolation.py:15-17.
__global__ void __launch_bounds__(kInterpBlock)
interp_weights_kernel(int n, int k, const float *__restrict__ dist2, float *__restrict__ weight) {
    const int row = blockIdx.x * kInterpBlock + threadIdx.x;
    if (row >= n) return;
    const float *d = dist2 + (size_t)row * k;
    float sum = 0.f;
    for (int i = 0; i < k; ++i) {
        float r = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(d + i)), 1e-8f));
        sum = __fadd_rn(sum, r);
    }
    for (int i = 0; i < k; ++i) {
        float r = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(d + i)), 1e-8f));
        weight[(size_t)row * k + i] = __fdiv_rn(r, sum);
    }
}

// One thread per (fine row, chunk); the thread keeps its chunk and walks rows (col_walk: no division per item),
// and the three (idx, weight) pairs of a row are requested before the first gather.
template <int VEC>
__global__ void __launch_bounds__(kInterpBlock)
interp_forward_kernel(long long n, int chunks, int c, int k, int m, const float *__restrict__ input,
                      const int *__restrict__ idx, const float *__restrict__ weight,
                      float *__restrict__ output) {
    const ColWalk cw = col_walk(chunks, kInterpBlock);
    for (long long row = cw.row; row < n; row += cw.row_step) {
        Chunk<VEC> acc = Chunk<VEC>::zero();
        if (k == 3) {
            int j[3];
            float w[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) { j[i] = __ldg(idx + row * 3 + i); w[i] = __ldg(weight + row * 3 + i); }
            Chunk<VEC> v[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                if (j[i] < 0) j[i] += m;  // python negative index (interpolation.py:21): no -1 masking
                v[i] = Chunk<VEC>::gather(input + (size_t)j[i] * c + cw.col * VEC);
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                v[i].scale(w[i]);   // separate multiply and add, as `new_feat += feat[idx] * weight` does
                acc.add(v[i]);
            }
        } else {
            for (int i = 0; i < k; ++i) {
                int j = __ldg(idx + row * k + i);
                if (j < 0) j += m;
                const float w = __ldg(weight + row * k + i);
                Chunk<VEC> v = Chunk<VEC>::gather(input + (size_t)j * c + cw.col * VEC);
                v.scale(w);
                acc.add(v);
            }
        }
        acc.store_stream(output + (size_t)row * c + cw.col * VEC);
    }
}

template <int VEC>
__global__ void __launch_bounds__(kInterpBlock)
interp_backward_kernel(long long m, int chunks, int c, int k, const float *__restrict__ grad_output,
                       const float *__restrict__ weight, const int *__restrict__ rowptr,
                       const int *__restrict__ perm, float *__restrict__ grad_input) {
    const long long total = m * chunks;
    const long long step = (long long)gridDim.x * kInterpBlock;
    for (long long t = (long long)blockIdx.x * kInterpBlock + threadIdx.x; t < total; t += step) {
        RowCol rc = split(t, chunks);
        Chunk<VEC> acc = Chunk<VEC>::zero();
        const int e_end = __ldg(rowptr + rc.row + 1);
#pragma unroll 4
        for (int e = __ldg(rowptr + rc.row); e < e_end; ++e) {
            const int p = __ldg(perm + e);
            const int q = p / k;
            const float w = __ldg(weight + p);
            acc.fma(Chunk<VEC>::gather(grad_output + (size_t)q * c + rc.col * VEC), w);
        }
        acc.store(grad_input + (size_t)rc.row * c + rc.col * VEC);
    }
}

// csr_walk.cuh policy: entry p = flat (fine row, slot); weight[p] * grad_output[p / k].
// KC = compile-time k (3: the interpolation of the model — p / 3 becomes a multiply-high instead of a runtime
// integer division per entry), 0 = runtime k.
template <int KC>
struct InterpBwdPolicy {
    const float *grad_output, *wgt;
    int c, k;
    static constexpr bool kWeighted = true;
    __device__ __forceinline__ float4 load(int p, int ch) const {
        const int q = KC > 0 ? p / KC : p / k;
        return ldg_gather4(grad_output + (size_t)q * c + ch * 4);
    }
    __device__ __forceinline__ float weight(int p, int) const { return __ldg(wgt + p); }
};

template <int KC>
static void launch_interp_walk(int m, int chunks, int c, int k, const float *grad_output, const float *weight,
                               const int *rowptr, const int *perm, float *grad_input, cudaStream_t st) {
    const InterpBwdPolicy<KC> pol{grad_output, weight, c, k};
    if (walk_batch() == 4)
        csr_walk_kernel<4, InterpBwdPolicy<KC>><<<walk_grid(m, chunks, 12), kWalkBlock, 0, st>>>(
            m, chunks, c, rowptr, perm, pol, 1.f, grad_input);
    else
        csr_walk_kernel<8, InterpBwdPolicy<KC>><<<walk_grid(m, chunks, 12), kWalkBlock, 0, st>>>(
            m, chunks, c, rowptr, perm, pol, 1.f, grad_input);
}

}  // namespace aopt

using namespace aopt;

extern "C" int aopt_interp_weights(int n, int k, const float *dist2, float *weight, aopt_stream_t stream) {
    if (n < 0 || k < 1 || k > AOPT_MAX_NSAMPLE) return AOPT_ERR_INVALID_ARGUMENT;
    if (n == 0) return AOPT_OK;
    if (!dist2 || !weight) return AOPT_ERR_INVALID_ARGUMENT;
    interp_weights_kernel<<<div_up(n, kInterpBlock), kInterpBlock, 0, as_stream(stream)>>>(n, k, dist2, weight);
    return check_launch();
}

extern "C" int aopt_interpolation_forward(int n, int c, int k, int m, const float *input, const int *idx,
                                          const float *weight, float *output, aopt_stream_t stream) {
    if (n < 0 || c < 1 || k < 1 || m < 1) return AOPT_ERR_INVALID_ARGUMENT;
    if (n == 0) return AOPT_OK;
    if (!input || !idx || !weight || !output) return AOPT_ERR_INVALID_ARGUMENT;
    const bool vec = (c % 4 == 0) && aligned16(input) && aligned16(output);
    if (vec) {
        const int chunks = c / 4;
        interp_forward_kernel<4><<<col_grid(n, chunks, kInterpBlock, 8), kInterpBlock, 0, as_stream(stream)>>>(
            n, chunks, c, k, m, input, idx, weight, output);
    } else {
        interp_forward_kernel<1><<<col_grid(n, c, kInterpBlock, 8), kInterpBlock, 0, as_stream(stream)>>>(
            n, c, c, k, m, input, idx, weight, output);
    }
    return check_launch();
}

extern "C" int aopt_interpolation_backward(int m, int c, int k, const float *grad_output, const float *weight,
                                           const int *rowptr, const int *perm, float *grad_input,
                                           aopt_stream_t stream) {
    if (m < 0 || c < 1 || k < 1) return AOPT_ERR_INVALID_ARGUMENT;
    if (m == 0) return AOPT_OK;
    if (!grad_output || !weight || !rowptr || !perm || !grad_input) return AOPT_ERR_INVALID_ARGUMENT;
    const bool vec = (c % 4 == 0) && aligned16(grad_output) && aligned16(grad_input);
    if (vec && use_batched_walk()) {
        const int chunks = c / 4;
        if (k == 3) launch_interp_walk<3>(m, chunks, c, k, grad_output, weight, rowptr, perm, grad_input, as_stream(stream));
        else launch_interp_walk<0>(m, chunks, c, k, grad_output, weight, rowptr, perm, grad_input, as_stream(stream));
    } else if (vec) {
        const int chunks = c / 4;
        interp_backward_kernel<4><<<stride_grid((long long)m * chunks, kInterpBlock, 8), kInterpBlock, 0, as_stream(stream)>>>(
            m, chunks, c, k, grad_output, weight, rowptr, perm, grad_input);
    } else {
        interp_backward_kernel<1><<<stride_grid((long long)m * c, kInterpBlock, 8), kInterpBlock, 0, as_stream(stream)>>>(
            m, c, c, k, grad_output, weight, rowptr, perm, grad_input);
    }
    return check_launch();
}
