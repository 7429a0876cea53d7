// legacy.cu — PTv1-layout fused ops kept so that every name of the reference `pointops` API on
// this path resolves.  Not called by PTv2m2.
//
//   aggregation  /root/reference/libs/pointops/src/aggregation/aggregation_cuda_kernel.cu:5-39
//                out[n,ch] = sum_s (input[idx[n,s],ch] + position[n,s,ch]) * weight[n,s,ch % w_c]
//                ("share-planes" layout; the PTv2 contiguous-group variant is gva.cu)
//   subtraction  /root/reference/libs/pointops/src/subtraction/subtraction_cuda_kernel.cu:5-30
//                out[n,s,ch] = input1[n,ch] - input2[idx[n,s],ch]
// Backward passes are atomic-free: per-query terms are plain stores, grad of the gathered input
// goes through the CSR (csr.cu).  idx < 0 (never produced for these ops by the reference callers)
// is treated as a zero row.
#include "common.cuh"

namespace aopt {

constexpr int kLegacyBlock = 256;

__global__ void __launch_bounds__(kLegacyBlock)
aggregation_forward_kernel(long long n, int k, int c, int w_c, const float *__restrict__ input,
                           const float *__restrict__ position, const float *__restrict__ weight,
                           const int *__restrict__ idx, float *__restrict__ output) {
    const long long total = n * c;
    const long long step = (long long)gridDim.x * kLegacyBlock;
    for (long long t = (long long)blockIdx.x * kLegacyBlock + threadIdx.x; t < total; t += step) {
        const long long pt = t / c;
        const int ch = (int)(t - pt * c);
        const int wc = ch % w_c;
        float acc = 0.f;
        for (int s = 0; s < k; ++s) {
            const int j = __ldg(idx + pt * k + s);
            const float in = j >= 0 ? __ldg(input + (size_t)j * c + ch) : 0.f;
            const float pos = __ldg(position + ((size_t)pt * k + s) * c + ch);
            const float w = __ldg(weight + ((size_t)pt * k + s) * w_c + wc);
            acc = fmaf(in + pos, w, acc);
        }
        output[t] = acc;
    }
}

// grad_position[n,s,ch] = go[n,ch] * w[n,s,ch % w_c]
__global__ void __launch_bounds__(kLegacyBlock)
aggregation_grad_position_kernel(long long n, int k, int c, int w_c, const float *__restrict__ weight,
                                 const float *__restrict__ grad_output, float *__restrict__ grad_position) {
    const long long total = n * k * c;
    const long long step = (long long)gridDim.x * kLegacyBlock;
    for (long long t = (long long)blockIdx.x * kLegacyBlock + threadIdx.x; t < total; t += step) {
        const long long row = t / c;  // (n,s)
        const int ch = (int)(t - row * c);
        const long long pt = row / k;
        grad_position[t] = __ldg(grad_output + pt * c + ch) * __ldg(weight + row * w_c + ch % w_c);
    }
}

// grad_weight[n,s,wc] = sum_{ch % w_c == wc} go[n,ch] * (input[idx[n,s],ch] + position[n,s,ch])
__global__ void __launch_bounds__(kLegacyBlock)
aggregation_grad_weight_kernel(long long n, int k, int c, int w_c, const float *__restrict__ input,
                               const float *__restrict__ position, const int *__restrict__ idx,
                               const float *__restrict__ grad_output, float *__restrict__ grad_weight) {
    const long long total = n * k * w_c;
    const long long step = (long long)gridDim.x * kLegacyBlock;
    for (long long t = (long long)blockIdx.x * kLegacyBlock + threadIdx.x; t < total; t += step) {
        const long long row = t / w_c;  // (n,s)
        const int wc = (int)(t - row * w_c);
        const long long pt = row / k;
        const int j = __ldg(idx + row);
        float acc = 0.f;
        for (int ch = wc; ch < c; ch += w_c) {
            const float in = j >= 0 ? __ldg(input + (size_t)j * c + ch) : 0.f;
            acc = fmaf(__ldg(grad_output + pt * c + ch), in + __ldg(position + row * c + ch), acc);
        }
        grad_weight[t] = acc;
    }
}

// grad_input[j,ch] = sum over CSR row j of go[q,ch] * w[p, ch % w_c],  q = p / k
__global__ void __launch_bounds__(kLegacyBlock)
aggregation_grad_input_kernel(long long n, int k, int c, int w_c, const float *__restrict__ weight,
                              const float *__restrict__ grad_output, const int *__restrict__ rowptr,
                              const int *__restrict__ perm, float *__restrict__ grad_input) {
    const long long total = n * c;
    const long long step = (long long)gridDim.x * kLegacyBlock;
    for (long long t = (long long)blockIdx.x * kLegacyBlock + threadIdx.x; t < total; t += step) {
        const long long j = t / c;
        const int ch = (int)(t - j * c);
        const int wc = ch % w_c;
        float acc = 0.f;
        const int e_end = __ldg(rowptr + j + 1);
        for (int e = __ldg(rowptr + j); e < e_end; ++e) {
            const int p = __ldg(perm + e);
            acc = fmaf(__ldg(grad_output + (size_t)(p / k) * c + ch), __ldg(weight + (size_t)p * w_c + wc), acc);
        }
        grad_input[t] = acc;
    }
}

template <int VEC>
__global__ void __launch_bounds__(kLegacyBlock)
subtraction_forward_kernel(long long rows, int k, int chunks, int c, const float *__restrict__ input1,
                           const float *__restrict__ input2, const int *__restrict__ idx,
                           float *__restrict__ out) {
    const long long total = rows * chunks;
    const long long step = (long long)gridDim.x * kLegacyBlock;
    for (long long t = (long long)blockIdx.x * kLegacyBlock + threadIdx.x; t < total; t += step) {
        RowCol rc = split(t, chunks);
        const int j = __ldg(idx + rc.row);
        Chunk<VEC> v = Chunk<VEC>::gather(input1 + (size_t)(rc.row / k) * c + rc.col * VEC);
        if (j >= 0) v.sub(Chunk<VEC>::gather(input2 + (size_t)j * c + rc.col * VEC));
        v.store_stream(out + (size_t)rc.row * c + rc.col * VEC);
    }
}

}  // namespace aopt

using namespace aopt;

extern "C" int aopt_aggregation_forward(int n, int nsample, int c, int w_c, const float *input,
                                        const float *position, const float *weight, const int *idx,
                                        float *output, aopt_stream_t stream) {
    if (n < 0 || nsample < 1 || c < 1 || w_c < 1) return AOPT_ERR_INVALID_ARGUMENT;
    if (n == 0) return AOPT_OK;
    if (!input || !position || !weight || !idx || !output) return AOPT_ERR_INVALID_ARGUMENT;
    aggregation_forward_kernel<<<stride_grid((long long)n * c, kLegacyBlock, 8), kLegacyBlock, 0, as_stream(stream)>>>(
        n, nsample, c, w_c, input, position, weight, idx, output);
    return check_launch();
}

extern "C" int aopt_aggregation_backward(int n, int nsample, int c, int w_c, const float *input,
                                         const float *position, const float *weight, const int *idx,
                                         const int *rowptr, const int *perm, const float *grad_output,
                                         float *grad_input, float *grad_position, float *grad_weight,
                                         aopt_stream_t stream) {
    if (n < 0 || nsample < 1 || c < 1 || w_c < 1) return AOPT_ERR_INVALID_ARGUMENT;
    if (n == 0) return AOPT_OK;
    if (!input || !position || !weight || !idx || !rowptr || !perm || !grad_output || !grad_input ||
        !grad_position || !grad_weight)
        return AOPT_ERR_INVALID_ARGUMENT;
    cudaStream_t st = as_stream(stream);
    aggregation_grad_position_kernel<<<stride_grid((long long)n * nsample * c, kLegacyBlock, 8), kLegacyBlock, 0, st>>>(
        n, nsample, c, w_c, weight, grad_output, grad_position);
    aggregation_grad_weight_kernel<<<stride_grid((long long)n * nsample * w_c, kLegacyBlock, 8), kLegacyBlock, 0, st>>>(
        n, nsample, c, w_c, input, position, idx, grad_output, grad_weight);
    aggregation_grad_input_kernel<<<stride_grid((long long)n * c, kLegacyBlock, 8), kLegacyBlock, 0, st>>>(
        n, nsample, c, w_c, weight, grad_output, rowptr, perm, grad_input);
    return check_launch(3);
}

extern "C" int aopt_subtraction_forward(int n, int nsample, int c, const float *input1, const float *input2,
                                        const int *idx, float *output, aopt_stream_t stream) {
    if (n < 0 || nsample < 1 || c < 1) return AOPT_ERR_INVALID_ARGUMENT;
    const long long rows = (long long)n * nsample;
    if (rows == 0) return AOPT_OK;
    if (!input1 || !input2 || !idx || !output) return AOPT_ERR_INVALID_ARGUMENT;
    const bool vec = (c % 4 == 0) && aligned16(input1) && aligned16(input2) && aligned16(output);
    if (vec) {
        const int chunks = c / 4;
        subtraction_forward_kernel<4><<<stride_grid(rows * chunks, kLegacyBlock, 8), kLegacyBlock, 0, as_stream(stream)>>>(
            rows, nsample, chunks, c, input1, input2, idx, output);
    } else {
        subtraction_forward_kernel<1><<<stride_grid(rows * c, kLegacyBlock, 8), kLegacyBlock, 0, as_stream(stream)>>>(
            rows, nsample, c, c, input1, input2, idx, output);
    }
    return check_launch();
}
