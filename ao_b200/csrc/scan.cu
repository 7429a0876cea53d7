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
    pdl_wait();      // no-op unless launched as a programmatic dependent (launch_exclusive_scan_chained)
    pdl_trigger();
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
    // Look-back by warp 0, 32 predecessors at a time (a single thread walking back one tile per L2 round trip made
    // the scan of 1.3 M cell counters take 23 us and every small scan 8-13 us: profiles/r02c_ops_L0_ncu_summary.md).
    // Lane l polls tile (base - l) until it is published; the nearest tile that already holds an inclusive prefix
    // ends the walk, the aggregates in front of it are summed.
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        volatile unsigned long long *st = state;
        int prefix = 0;
        if (tile == 0) {
            if (lane == 0) st[0] = kFlagInclusive | (unsigned)total;
        } else {
            if (lane == 0) st[tile] = kFlagAggregate | (unsigned)total;
            int base = tile - 1;
            for (;;) {
                const int p = base - lane;
                unsigned long long sv = kFlagInclusive;             // tiles before the first: inclusive prefix 0
                if (p >= 0) { do { sv = st[p]; } while ((sv >> 32) == 0); }
                const bool inclusive = (sv >> 32) == 2;
                const unsigned inc_mask = __ballot_sync(0xffffffffu, inclusive);
                const int first = inc_mask ? __ffs(inc_mask) - 1 : 31;   // nearest predecessor with an inclusive prefix
                int v = lane <= first ? (int)(unsigned)(sv & 0xffffffffull) : 0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                prefix += v;
                if (inc_mask) break;
                base -= 32;
            }
            if (lane == 0) st[tile] = kFlagInclusive | (unsigned)(prefix + total);
        }
        if (lane == 0) prefix_s = prefix;
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

static void launch_scan(const int *in, int *out, int n, int *partial, cudaStream_t st, bool zero) {
    const int tiles = div_up(n, kScanTile);
    if (tiles == 0) {
        scan_empty_kernel<<<1, 1, 0, st>>>(out);
        return;
    }
    unsigned long long *state = reinterpret_cast<unsigned long long *>(partial);
    unsigned *counter = reinterpret_cast<unsigned *>(state + tiles + 1);
    if (zero) cudaMemsetAsync(partial, 0, sizeof(int) * scan_partial_ints(n), st);
    scan_onepass_kernel<<<tiles, kScanBlock, 0, st>>>(n, in, out, state, counter);
}

void launch_exclusive_scan(const int *in, int *out, int n, int *partial, cudaStream_t st) { launch_scan(in, out, n, partial, st, true); }
void launch_exclusive_scan_chained(const int *in, int *out, int n, int *partial, cudaStream_t st, bool pdl) {
    const int tiles = div_up(n, kScanTile);
    if (tiles == 0) {
        scan_empty_kernel<<<1, 1, 0, st>>>(out);
        return;
    }
    unsigned long long *state = reinterpret_cast<unsigned long long *>(partial);
    unsigned *counter = reinterpret_cast<unsigned *>(state + tiles + 1);
    launch_chain(pdl, scan_onepass_kernel, tiles, kScanBlock, 0, st, n, in, out, state, counter);
}
void launch_exclusive_scan_prezeroed(const int *in, int *out, int n, int *partial, cudaStream_t st) {
    launch_scan(in, out, n, partial, st, false);
}

}  // namespace aopt
