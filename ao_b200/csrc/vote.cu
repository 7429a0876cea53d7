// vote.cu — fragment vote of the tester: pred[index[r], :] += softmax(logits[r, :]).
//
// Replaces the torch ops of /root/reference/pointcept/engines/test.py:106-113
//   pred_part = F.softmax(pred_part, -1)                      (n, classes) materialised
//   for be in offset: pred[idx_part[bs:be], :] += pred_part[bs:be]   gather + add + index_put per scene
// (softmax kernel + three indexing kernels per scene, and an int64 index temp) with one kernel per scene range:
// the class probabilities never reach memory.  The rows of ONE range carry distinct indices (a fragment holds one
// point per voxel, pointcept/datasets/transform.py:834-837), so the read-modify-write needs no atomics and the
// result is deterministic; successive ranges — which may vote for the same point again — are successive launches
// on the caller's stream, in the reference's order.
//
// Thread = one row; the row (13-20 classes, 52-80 bytes) is walked three times (max, sum of exp, write), the second
// and third pass out of L1.  fp32 arithmetic in torch's order: exp(x - max) with expf, sequential sum, IEEE divide.
#include "common.cuh"

namespace aopt {

constexpr int kVoteBlock = 128;

__global__ void __launch_bounds__(kVoteBlock)
vote_accumulate_kernel(int rows, int c, long long n_pred, const float *__restrict__ logits,
                       const long long *__restrict__ index, float *__restrict__ pred, int *__restrict__ bad) {
    const int r = blockIdx.x * kVoteBlock + threadIdx.x;
    if (r >= rows) return;
    long long j = index[r];
    if (j < 0) j += n_pred;                                   // python negative index
    if (j < 0 || j >= n_pred) { if (bad) atomicExch(bad, 1); return; }
    const float *x = logits + (size_t)r * c;
    float mx = __ldg(x);
    for (int i = 1; i < c; ++i) mx = fmaxf(mx, __ldg(x + i));
    float sum = 0.f;
    for (int i = 0; i < c; ++i) sum = __fadd_rn(sum, expf(__fsub_rn(__ldg(x + i), mx)));
    float *p = pred + (size_t)j * c;
    for (int i = 0; i < c; ++i) p[i] = __fadd_rn(p[i], __fdiv_rn(expf(__fsub_rn(__ldg(x + i), mx)), sum));
}

}  // namespace aopt

using namespace aopt;

extern "C" int aopt_vote_accumulate(int rows, int c, long long n_pred, const float *logits, const long long *index,
                                    float *pred, int *bad_flag, aopt_stream_t stream) {
    if (rows < 0 || c < 1 || n_pred < 0) return AOPT_ERR_INVALID_ARGUMENT;
    if (rows == 0) return AOPT_OK;
    if (!logits || !index || !pred || n_pred == 0) return AOPT_ERR_INVALID_ARGUMENT;
    vote_accumulate_kernel<<<div_up(rows, kVoteBlock), kVoteBlock, 0, as_stream(stream)>>>(rows, c, n_pred, logits, index,
                                                                                         pred, bad_flag);
    return check_launch();
}
