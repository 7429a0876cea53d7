// fps.cu — farthest point sampling over the offset-encoded batch (the PTv1 caller,
// /root/reference/pointcept/models/point_transformer/point_transformer_seg.py:101).
//
// Replaces /root/reference/libs/pointops/src/sampling/sampling_cuda_kernel.cu:14-122: ONE CTA per scene
// (4 of the 148 SMs busy for an S3DIS batch), every one of the m iterations re-reading xyz and the
// running distance `tmp` of the whole scene from global memory, then a 10-step shared-memory tree.
//
// Here a thread-block CLUSTER owns a scene (8 CTAs, 16 when the scene needs it):
//   * the scene's points and their running minimum distance live in REGISTERS for the whole kernel
//     (PPT <= 20 points per thread: 8 x 512 x 20 = 82k points per cluster of 8, 164k per cluster of 16), so an iteration touches no
//     global memory except the 4-byte result;
//   * per iteration: register update + local arg-max, two redux.sync per warp, one __syncthreads for
//     the CTA, then the CTA's winner (rank + coordinates) is stored into every peer CTA's shared
//     memory (DSMEM) and ONE cluster barrier publishes it; every warp then reads the cluster's slots
//     locally — the next centre never goes through L2.
// Scenes that do not fit the registers of a 16-CTA cluster use the same kernel with the points left
// in global memory (PPT = 0).
//
// Bit-compatibility with the reference.  Distance: SASS of the reference kernel built for sm_100a is
// FADD dy; FADD dx; FMUL dy*dy; FADD dz; FFMA dx*dx+.; FFMA dz*dz+. — dist2_ref().  tmp = min(d, tmp)
// (FMNMX).  Arg-max ties: the reference thread `tid` scans k = start+tid, start+tid+B, ... keeping the
// first strict maximum, and its tree keeps the LOWER tid on equal values (`v2 > v1 ? i2 : i1`), B =
// opt_n_threads(n_max) (cuda_utils.h:11-14).  The tree merges slot s with slot s + h for h = B/2 ... 1,
// so two tied threads meet at the level of their lowest differing tid bit and the one with a 0 there
// wins: among equal maxima the winner is the one with the smallest (bitrev(tid), k), tid = (k - start)
// mod B.  The rank (dist bits + 1, ~(bitrev(tid) << 22 | k - start)) encodes exactly that order, so the
// result is the reference's for every input, ties included (scenes below 2^22 points).
#include <cooperative_groups.h>

#include "common.cuh"
#include "knn_common.cuh"

namespace cg = cooperative_groups;

namespace aopt {

constexpr int kFpsThreads = 512;
constexpr int kFpsMaxCluster = 16;
constexpr int kFpsMaxPpt = 20;  // points per thread kept in registers (24 spills under the 128-register cap of 512 threads)

// A candidate is ranked by (hi, lo): hi = bits of its running distance + 1 (0 = "no point"; distances are
// >= 0 so the bit pattern orders like the value), lo = ~tie with
// tie = bitrev(reference thread) << 22 | index in the scene — the order the reference's tree resolves equal
// distances in (see the header).
struct __align__(16) FpsSlot {
    unsigned hi, lo;
    float x, y, z;
    int pad[3];
};

__device__ __forceinline__ unsigned fps_tie_lo(int local, int tie_mask, int rev_shift) {
    const unsigned rt = (unsigned)(((unsigned long long)__brev((unsigned)(local & tie_mask))) >> rev_shift);
    return 0xffffffffu - ((rt << 22) | ((unsigned)local & 0x3fffffu));
}

// Arg-max of (hi, lo) over the lanes named by `mask`-less full warp; returns the winning lane.
__device__ __forceinline__ int fps_warp_argmax(unsigned hi, unsigned lo, unsigned &mh, unsigned &ml) {
    mh = __reduce_max_sync(0xffffffffu, hi);
    ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    const unsigned who = __ballot_sync(0xffffffffu, hi == mh && (lo == ml || mh == 0u));
    return __ffs(who) - 1;
}

template <int PPT>
__global__ void __launch_bounds__(kFpsThreads, 1)
fps_cluster_kernel(int b, int tie_mask, int rev_shift, const float *__restrict__ xyz,
                   const int *__restrict__ offset, const int *__restrict__ new_offset, float *__restrict__ tmp,
                   int *__restrict__ idx) {
    cg::cluster_group cluster = cg::this_cluster();
    const int cl = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int scene = blockIdx.x / cl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int start_n = scene == 0 ? 0 : __ldg(offset + scene - 1), end_n = __ldg(offset + scene);
    const int start_m = scene == 0 ? 0 : __ldg(new_offset + scene - 1), end_m = __ldg(new_offset + scene);
    const int n = end_n - start_n, m = end_m - start_m;
    if (m <= 0) return;  // the whole cluster leaves together

    __shared__ FpsSlot warp_slot[kFpsThreads / 32];
    __shared__ FpsSlot slots[2][kFpsMaxCluster];

    if (rank == 0 && tid == 0) idx[start_m] = start_n;  // sampling_cuda_kernel.cu:38
    // Point p of this thread has scene-local index first + p * stride.  stride = cl * 512 is a multiple of
    // the reference block size (the launcher guarantees it), so all points of a thread belong to ONE
    // reference thread and their tie order is the scan order p = 0, 1, ...: a strict '>' in the loop is the
    // reference's in-thread rule, and the (hi, lo) rank is only built once per iteration.
    const int stride = cl * kFpsThreads;
    const int first = rank * kFpsThreads + tid;

    float px[PPT > 0 ? PPT : 1], py[PPT > 0 ? PPT : 1], pz[PPT > 0 ? PPT : 1], pt[PPT > 0 ? PPT : 1];
    if (PPT > 0) {
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const int loc = first + p * stride;
            const bool ok = loc < n;
            const size_t g = (size_t)(start_n + (ok ? loc : 0));
            px[p] = n > 0 ? __ldg(xyz + g * 3) : 0.f;
            py[p] = n > 0 ? __ldg(xyz + g * 3 + 1) : 0.f;
            pz[p] = n > 0 ? __ldg(xyz + g * 3 + 2) : 0.f;
            pt[p] = (ok && n > 0) ? tmp[g] : -1.f;  // -1 stays -1 under min() and never beats `best = -1`
        }
    }
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (n > 0) {
        cx = __ldg(xyz + (size_t)start_n * 3); cy = __ldg(xyz + (size_t)start_n * 3 + 1);
        cz = __ldg(xyz + (size_t)start_n * 3 + 2);
    }

    for (int j = 1; j < m; ++j) {
        float bt = -1.f, bx = 0.f, by = 0.f, bz = 0.f;  // sampling_cuda_kernel.cu:43-44
        int bloc = 0;
        if (PPT > 0) {
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                const float t = fminf(dist2_ref(px[p], py[p], pz[p], cx, cy, cz), pt[p]);
                pt[p] = t;
                const bool up = t > bt;
                bt = up ? t : bt;
                bloc = up ? p : bloc;
                bx = up ? px[p] : bx; by = up ? py[p] : by; bz = up ? pz[p] : bz;
            }
            bloc = first + bloc * stride;
        } else {
            for (int loc = first; loc < n; loc += stride) {
                const size_t g = (size_t)(start_n + loc);
                const float x = __ldg(xyz + g * 3), y = __ldg(xyz + g * 3 + 1), z = __ldg(xyz + g * 3 + 2);
                const float t = fminf(dist2_ref(x, y, z, cx, cy, cz), tmp[g]);
                tmp[g] = t;
                const bool up = t > bt;
                bt = up ? t : bt;
                bloc = up ? loc : bloc;
                bx = up ? x : bx; by = up ? y : by; bz = up ? z : bz;
            }
        }
        const unsigned hi = bt < 0.f ? 0u : __float_as_uint(bt) + 1u;
        const unsigned lo = fps_tie_lo(bloc, tie_mask, rev_shift);
        unsigned mh, ml;
        const int wl = fps_warp_argmax(hi, lo, mh, ml);
        if (lane == wl) {
            FpsSlot s; s.hi = mh; s.lo = ml; s.x = bx; s.y = by; s.z = bz; s.pad[0] = s.pad[1] = s.pad[2] = 0;
            warp_slot[warp] = s;
        }
        __syncthreads();
        // CTA winner -> slot `rank` of every CTA in the cluster (DSMEM), double-buffered by iteration parity
        if (warp == 0) {
            FpsSlot s = warp_slot[lane & (kFpsThreads / 32 - 1)];
            if (lane >= kFpsThreads / 32) s.hi = 0u, s.lo = 0u;
            const int cw = fps_warp_argmax(s.hi, s.lo, mh, ml);
            s.hi = mh; s.lo = ml;
            s.x = __shfl_sync(0xffffffffu, s.x, cw); s.y = __shfl_sync(0xffffffffu, s.y, cw);
            s.z = __shfl_sync(0xffffffffu, s.z, cw);
            if (lane < cl) *cluster.map_shared_rank(&slots[j & 1][rank], lane) = s;
        }
        cluster.sync();  // barrier.cluster arrive.release / wait.acquire: the remote stores are visible
        {
            FpsSlot s = slots[j & 1][lane & (kFpsMaxCluster - 1)];
            if (lane >= cl) s.hi = 0u, s.lo = 0u;
            const int cw = fps_warp_argmax(s.hi, s.lo, mh, ml);
            cx = __shfl_sync(0xffffffffu, s.x, cw); cy = __shfl_sync(0xffffffffu, s.y, cw);
            cz = __shfl_sync(0xffffffffu, s.z, cw);
        }
        if (rank == 0 && tid == 0)
            idx[start_m + j] = start_n + (mh != 0u ? (int)((0xffffffffu - ml) & 0x3fffffu) : 0);
    }
    if (PPT > 0) {  // the reference leaves the final running distances in tmp
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const int loc = first + p * stride;
            if (loc < n) tmp[(size_t)(start_n + loc)] = pt[p];
        }
    }
    cluster.sync();  // no CTA exits while a peer may still store into its shared memory
}

template <int PPT>
static cudaError_t launch_fps(int b, int cl, int tie_mask, int rev_shift, const float *xyz, const int *offset,
                              const int *new_offset, float *tmp, int *idx, cudaStream_t st) {
    auto kern = fps_cluster_kernel<PPT>;
    if (cl > 8) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(b * cl), 1, 1);
    cfg.blockDim = dim3(kFpsThreads, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, b, tie_mask, rev_shift, xyz, offset, new_offset, tmp, idx);
}

}  // namespace aopt

using namespace aopt;

// Same arguments as farthest_point_sampling_cuda_launcher (sampling_cuda_kernel.h): b scenes, n_max =
// the largest scene (it selects the reference's block size and with it the tie order), xyz (n,3),
// offset / new_offset (b) cumulative ends, tmp (n) running distances pre-filled with 1e10 by the caller
// (sampling.py:19), idx (new_offset[b-1]) output.  Identifiers beyond 2^22 points per scene are out of
// the key's tie field: AOPT_ERR_UNSUPPORTED.
extern "C" int aopt_farthest_point_sampling(int b, int n_max, const float *xyz, const int *offset,
                                            const int *new_offset, float *tmp, int *idx,
                                            aopt_stream_t stream) {
    if (b < 0 || n_max < 0) return AOPT_ERR_INVALID_ARGUMENT;
    if (b == 0) return AOPT_OK;
    if (!xyz || !offset || !new_offset || !tmp || !idx) return AOPT_ERR_INVALID_ARGUMENT;
    if (n_max >= (1 << 22)) return AOPT_ERR_UNSUPPORTED;
    // reference block size: opt_n_threads(n_max) = min(2^floor(log2 n_max), 1024)   (cuda_utils.h:11-14)
    int block_ref = 1, log2_ref = 0;
    while (block_ref * 2 <= n_max && block_ref < 1024) { block_ref *= 2; ++log2_ref; }
    const int tie_mask = block_ref - 1, rev_shift = 32 - log2_ref;
    // cluster size: cl * 512 must be a multiple of block_ref (kernel comment) => cl >= 2 when block_ref = 1024
    int cl = 8;
    if (n_max > 8 * kFpsThreads * kFpsMaxPpt) cl = 16;
    else if (n_max <= 2 * kFpsThreads * 4) cl = block_ref <= kFpsThreads ? 1 : 2;  // tiny scenes: the cluster barrier would dominate
    if (const char *e = getenv("AOPT_FPS_CLUSTER")) {
        const int v = atoi(e);
        if ((v == 1 && block_ref <= kFpsThreads) || v == 2 || v == 4 || v == 8 || v == 16) cl = v;
    }
    const long long per_thread = ((long long)n_max + (long long)cl * kFpsThreads - 1) / ((long long)cl * kFpsThreads);
    cudaError_t e;
    cudaStream_t st = as_stream(stream);
    if (per_thread <= 4) e = launch_fps<4>(b, cl, tie_mask, rev_shift, xyz, offset, new_offset, tmp, idx, st);
    else if (per_thread <= 8) e = launch_fps<8>(b, cl, tie_mask, rev_shift, xyz, offset, new_offset, tmp, idx, st);
    else if (per_thread <= 12) e = launch_fps<12>(b, cl, tie_mask, rev_shift, xyz, offset, new_offset, tmp, idx, st);
    else if (per_thread <= 16) e = launch_fps<16>(b, cl, tie_mask, rev_shift, xyz, offset, new_offset, tmp, idx, st);
    else if (per_thread <= kFpsMaxPpt) e = launch_fps<kFpsMaxPpt>(b, cl, tie_mask, rev_shift, xyz, offset, new_offset, tmp, idx, st);
    else e = launch_fps<0>(b, cl, tie_mask, rev_shift, xyz, offset, new_offset, tmp, idx, st);
    if (e != cudaSuccess) return AOPT_ERR_LAUNCH;
    return check_launch();
}
