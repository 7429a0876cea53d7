// scan.cu — three-kernel exclusive scan (tile scan → scan of tile totals → add back).
#include "scan.cuh"

namespace aopt {

// Inclusive scan of one int per thread across a 1024-thread block; returns the inclusive value
// and leaves the block total in warp_sums[31].
__device__ __forceinline__ int block_inclusive_scan(int v, int *warp_sums) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = warp_sums[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += t;
        }
        warp_sums[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    return inc + (warp > 0 ? warp_sums[warp - 1] : 0);
}

// in/out may alias: every thread reads its own 4 elements before writing them.
__global__ void __launch_bounds__(kScanBlock)
scan_tiles_kernel(int n, const int *in, int *out, int *__restrict__ partial) {
    __shared__ int warp_sums[kScanBlock / 32];
    const long long base = (long long)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems];
    int sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        sum += v[i];
    }
    int excl = block_inclusive_scan(sum, warp_sums) - sum;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = excl;
        excl += v[i];
    }
    if (threadIdx.x == kScanBlock - 1) partial[blockIdx.x] = warp_sums[kScanBlock / 32 - 1];
}

// Single block: exclusive scan of the tile totals in place; partial[n_tiles] = grand total.
__global__ void __launch_bounds__(kScanBlock)
scan_partials_kernel(int n_tiles, int *__restrict__ partial) {
    __shared__ int warp_sums[kScanBlock / 32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += kScanBlock) {
        const int i = base + threadIdx.x;
        const int v = i < n_tiles ? partial[i] : 0;
        const int inc = block_inclusive_scan(v, warp_sums);
        const int carry = carry_s;
        if (i < n_tiles) partial[i] = carry + inc - v;
        __syncthreads();  // everyone has read carry_s and warp_sums
        if (threadIdx.x == kScanBlock - 1) carry_s = carry + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[n_tiles] = carry_s;
}

__global__ void __launch_bounds__(256)
scan_add_kernel(int n, int *__restrict__ out, const int *__restrict__ partial, int n_tiles) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i < n) out[i] += partial[i / kScanTile];
    if (i == n) out[n] = partial[n_tiles];
}

void launch_exclusive_scan(const int *in, int *out, int n, int *partial, cudaStream_t st) {
    const int tiles = div_up(n, kScanTile);
    if (tiles > 0) scan_tiles_kernel<<<tiles, kScanBlock, 0, st>>>(n, in, out, partial);
    scan_partials_kernel<<<1, kScanBlock, 0, st>>>(tiles, partial);
    scan_add_kernel<<<div_up((long long)n + 1, 256), 256, 0, st>>>(n, out, partial, tiles);
}

}  // namespace aopt
