// scan.cu — single-pass exclusive scan with decoupled look-back.
//
// Every CTA takes the next tile id from an atomic counter (so a tile's predecessors are always running
// or done), scans its 2048 elements, publishes its aggregate, then walks back over the published states
// of the preceding tiles until it meets an inclusive prefix.  A state is one 64-bit word
// (flag << 32 | value) written with a single store, so flag and value are always consistent.
#include "scan.cuh"

namespace aopt {

constexpr unsigned long long kFlagAggregate = 1ull << 32;
constexpr unsigned long long kFlagInclusive = 2ull << 32;

// Inclusive scan of one int per thread across the block; returns the inclusive value, block total in *total.
__device__ __forceinline__ int block_inclusive_scan(int v, int *warp_sums, int *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kScanBlock / 32;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < kWarps ? warp_sums[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += t;
        }
        if (lane < kWarps) warp_sums[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    *total = warp_sums[kWarps - 1];
    return inc + (warp > 0 ? warp_sums[warp - 1] : 0);
}

// in/out may alias: every thread reads its own elements before any element of the tile is written.
__global__ void __launch_bounds__(kScanBlock)
scan_onepass_kernel(int n, const int *in, int *out, unsigned long long *state, unsigned *counter) {
    __shared__ int warp_sums[kScanBlock / 32];
    __shared__ int tile_s, prefix_s;
    if (threadIdx.x == 0) tile_s = (int)atomicAdd(counter, 1u);
    __syncthreads();
    const int tile = tile_s;
    const long long base = (long long)tile * kScanTile + (long long)threadIdx.x * kScanItems;
    int v[kScanItems];
    int sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        sum += v[i];
    }
    int total;
    const int inc = block_inclusive_scan(sum, warp_sums, &total);
    if (threadIdx.x == 0) {
        volatile unsigned long long *st = state;
        int prefix = 0;
        if (tile == 0) {
            st[0] = kFlagInclusive | (unsigned)total;
        } else {
            st[tile] = kFlagAggregate | (unsigned)total;
            for (int p = tile - 1; p >= 0; --p) {
                unsigned long long s;
                do { s = st[p]; } while ((s >> 32) == 0);  // predecessor not published yet
                prefix += (int)(unsigned)(s & 0xffffffffull);
                if ((s >> 32) == 2) break;
            }
            st[tile] = kFlagInclusive | (unsigned)(prefix + total);
        }
        prefix_s = prefix;
    }
    __syncthreads();
    int excl = prefix_s + inc - sum;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = excl;
        excl += v[i];
    }
    // the tile that holds element n-1 also writes the grand total
    if (threadIdx.x == kScanBlock - 1 && (long long)(tile + 1) * kScanTile >= n) out[n] = prefix_s + total;
}

__global__ void scan_empty_kernel(int *out) { out[0] = 0; }

void launch_exclusive_scan(const int *in, int *out, int n, int *partial, cudaStream_t st) {
    const int tiles = div_up(n, kScanTile);
    if (tiles == 0) {
        scan_empty_kernel<<<1, 1, 0, st>>>(out);
        return;
    }
    unsigned long long *state = reinterpret_cast<unsigned long long *>(partial);
    unsigned *counter = reinterpret_cast<unsigned *>(state + tiles + 1);
    cudaMemsetAsync(partial, 0, sizeof(int) * scan_partial_ints(n), st);
    scan_onepass_kernel<<<tiles, kScanBlock, 0, st>>>(n, in, out, state, counter);
}

}  // namespace aopt
