// bbox.cu — see bbox.cuh.
#include "bbox.cuh"

namespace aopt {

constexpr int kBboxBlock = 256;
constexpr int kBboxPerThread = 8;
constexpr int kBboxChunk = kBboxBlock * kBboxPerThread;

template <bool WITH_MAX>
__global__ void __launch_bounds__(kBboxBlock)
scene_bbox_kernel(int n, int b, const float *__restrict__ xyz, const int *__restrict__ offset,
                  unsigned *__restrict__ lo, unsigned *__restrict__ hi) {
    __shared__ float red[6][kBboxBlock / 32];
    __shared__ int seg_s[2];
    pdl_wait();      // chained launch from the kNN grid build (no-op otherwise)
    pdl_trigger();
    const int base = blockIdx.x * kBboxChunk;
    const int last = min(base + kBboxChunk, n) - 1;
    if (threadIdx.x < 2) seg_s[threadIdx.x] = find_segment(threadIdx.x == 0 ? base : last, offset, b);
    __syncthreads();
    const int sc_first = seg_s[0], sc_last = seg_s[1];
    if (sc_first == sc_last) {
        // the whole chunk lies in one scene: block reduction, then six atomics
        if (sc_first >= b) return;  // past the last offset: belongs to no scene
        float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
        // unconditional loads (index clamped to the chunk's last point: duplicates do not change a
        // min/max), so all 24 requests of a thread are in flight together
        float v[kBboxPerThread][3];
#pragma unroll
        for (int u = 0; u < kBboxPerThread; ++u) {
            const int i = min(base + u * kBboxBlock + (int)threadIdx.x, last);
#pragma unroll
            for (int a = 0; a < 3; ++a) v[u][a] = __ldg(xyz + (size_t)i * 3 + a);
        }
#pragma unroll
        for (int u = 0; u < kBboxPerThread; ++u) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                mn[a] = fminf(mn[a], v[u][a]);
                if (WITH_MAX) mx[a] = fmaxf(mx[a], v[u][a]);
            }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], d));
                if (WITH_MAX) mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], d));
            }
            if ((threadIdx.x & 31) == 0) {
                red[a][threadIdx.x >> 5] = mn[a];
                if (WITH_MAX) red[3 + a][threadIdx.x >> 5] = mx[a];
            }
        }
        __syncthreads();
        if (threadIdx.x < (WITH_MAX ? 6 : 3)) {
            const int a = threadIdx.x;
            float v = red[a][0];
            for (int w = 1; w < kBboxBlock / 32; ++w) v = a < 3 ? fminf(v, red[a][w]) : fmaxf(v, red[a][w]);
            if (a < 3) atomicMin(lo + sc_first * 3 + a, bbox_encode(v));
            else atomicMax(hi + sc_first * 3 + (a - 3), bbox_encode(v));
        }
    } else {
        // chunk straddles scene boundaries (at most b-1 chunks).  Per-point atomics serialise on the
        // few (scene, axis) words — 2048 x 6 same-address atomics cost ~25 us per launch whatever n is
        // (profiles/r01i: scene_bbox<1> 27 us at 12k and at 320k points).  Instead: one masked warp
        // reduction per scene present in the chunk, six atomics per warp and scene.
        float v[kBboxPerThread][3];
        int pi[kBboxPerThread];
#pragma unroll
        for (int u = 0; u < kBboxPerThread; ++u) {
            pi[u] = base + u * kBboxBlock + (int)threadIdx.x;
            const int i = min(pi[u], last);
#pragma unroll
            for (int a = 0; a < 3; ++a) v[u][a] = __ldg(xyz + (size_t)i * 3 + a);
        }
        const int sc_end = min(sc_last, b - 1);
        for (int sc = sc_first; sc <= sc_end; ++sc) {
            const int s0 = sc == 0 ? 0 : __ldg(offset + sc - 1), s1 = __ldg(offset + sc);
            float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
            bool any = false;
#pragma unroll
            for (int u = 0; u < kBboxPerThread; ++u) {
                const bool in = pi[u] <= last && pi[u] >= s0 && pi[u] < s1;
                any |= in;
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    mn[a] = in ? fminf(mn[a], v[u][a]) : mn[a];
                    if (WITH_MAX) mx[a] = in ? fmaxf(mx[a], v[u][a]) : mx[a];
                }
            }
            if (!__any_sync(0xffffffffu, any)) continue;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], d));
                    if (WITH_MAX) mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], d));
                }
            }
            if ((threadIdx.x & 31) == 0) {
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    atomicMin(lo + sc * 3 + a, bbox_encode(mn[a]));
                    if (WITH_MAX) atomicMax(hi + sc * 3 + a, bbox_encode(mx[a]));
                }
            }
        }
    }
}

void launch_scene_bbox(int n, int b, const float *xyz, const int *offset, unsigned *lo, unsigned *hi,
                       cudaStream_t st, bool init, bool pdl) {
    if (init) {
        cudaMemsetAsync(lo, 0xff, sizeof(unsigned) * 3 * (size_t)b, st);
        if (hi) cudaMemsetAsync(hi, 0x00, sizeof(unsigned) * 3 * (size_t)b, st);
    }
    if (n <= 0) return;
    const int grid = div_up(n, kBboxChunk);
    if (hi) launch_chain(pdl, scene_bbox_kernel<true>, grid, kBboxBlock, 0, st, n, b, xyz, offset, lo, hi);
    else launch_chain(pdl, scene_bbox_kernel<false>, grid, kBboxBlock, 0, st, n, b, xyz, offset, lo, hi);
}

}  // namespace aopt
