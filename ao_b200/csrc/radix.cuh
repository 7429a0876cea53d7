// radix.cuh — stable LSD radix sort of (key, int32 value) pairs, 11-bit digits, keys of 32 or 64 bits.
//
// Two users, both of which used to lean on something slower:
//   * the GridPool voxel partition (pool.cu): points sorted by compact voxel key — was torch.sort on 64-bit keys
//     (eight library onesweep passes per stage; /root/reference/pointcept/models/point_transformer_v2/
//     point_transformer_v2m2_base.py:260-263 does torch.unique + torch.sort);
//   * the transposed neighbour graph (csr.cu): flat (query, slot) positions sorted by source index — was
//     count (atomics) + fill + an all-pairs rank inside every row, quadratic in the in-degree of hub rows.
// A stable sort gives both the order they need for free: ascending point id inside a voxel, ascending flat
// position inside a CSR row (the fixed summation order of every atomic-free backward pass).
//
// One pass = three launches on the caller's stream:
//   radix_hist_kernel      per-tile digit histogram  -> table[digit * tiles + tile]
//   launch_exclusive_scan  over the table (digit-major: all tiles of digit 0, then digit 1, ...)
//   radix_scatter_kernel   stable rank inside the tile + the scanned base -> final position of the pass
// A tile is 2048 consecutive keys handled by 8 warps; warp w owns keys [256 w, 256 w + 256) of the tile and ranks
// them 32 at a time with __match_any_sync against its private histogram, so the rank of a key is the number of
// keys with the same digit before it in the tile — stability without a second sort.
// `npass_dev` (may be NULL): device word holding the number of passes actually needed (the voxel key width is
// known only on the device); kernels of later passes return at once, the caller picks the buffer that holds the
// result (radix_result_in_second).
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace aopt {

constexpr int kRadixBits = 11;
constexpr int kRadixBins = 1 << kRadixBits;
constexpr int kRadixBlock = 256;
constexpr int kRadixWarps = kRadixBlock / 32;
constexpr int kRadixTile = 2048;
constexpr int kRadixRounds = kRadixTile / kRadixBlock;  // keys per thread

inline int radix_tiles(long long n) { return div_up(n, kRadixTile); }
// A CTA takes `tiles per block` consecutive tiles (one histogram column for all of them), so that the table the
// scan walks stays small next to the data: at most ~4 CTAs per SM.
inline int radix_tiles_per_block(long long n) { return div_up(radix_tiles(n), 4 * kNumSM) > 0 ? div_up(radix_tiles(n), 4 * kNumSM) : 1; }
inline int radix_blocks(long long n) { return div_up(radix_tiles(n), radix_tiles_per_block(n)); }
inline size_t radix_table_ints(long long n) { return (size_t)kRadixBins * radix_blocks(n) + 1; }
// scratch (ints) of one sort of n keys: the table + the scan's look-back states
inline size_t radix_scratch_ints(long long n) {
    return radix_table_ints(n) + 2 + scan_partial_ints((long long)radix_table_ints(n));
}

template <typename K>
struct PtrKeys {
    const K *p;
    __device__ __forceinline__ K operator()(int i) const { return p[i]; }
};

template <typename K, class Load>
__global__ void __launch_bounds__(kRadixBlock)
radix_hist_kernel(Load load, int n, int shift, int pass, const int *__restrict__ npass_dev, int tpb,
                  int *__restrict__ table, int *__restrict__ scan_state, int scan_state_ints) {
    pdl_wait();      // chained launches (launch_radix_pass with pdl): no-op otherwise
    pdl_trigger();
    // the look-back states / tile counter of this pass's scan (launch_exclusive_scan_prezeroed), also when the pass is
    // skipped: the scan kernel still runs
    for (int i = blockIdx.x * kRadixBlock + threadIdx.x; i < scan_state_ints; i += gridDim.x * kRadixBlock) scan_state[i] = 0;
    if (npass_dev && pass >= *npass_dev) return;
    __shared__ int hist[kRadixBins];
    for (int d = threadIdx.x; d < kRadixBins; d += kRadixBlock) hist[d] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * tpb * kRadixTile;
    const long long end = min(base + (long long)tpb * kRadixTile, (long long)n);
    for (long long i = base + threadIdx.x; i < end; i += kRadixBlock)
        atomicAdd(&hist[(int)((load((int)i) >> shift) & (K)(kRadixBins - 1))], 1);
    __syncthreads();
    for (int d = threadIdx.x; d < kRadixBins; d += kRadixBlock) table[(size_t)d * gridDim.x + blockIdx.x] = hist[d];
}

template <typename K, class Load>
__global__ void __launch_bounds__(kRadixBlock)
radix_scatter_kernel(Load load, const int *__restrict__ vin, K *__restrict__ kout, int *__restrict__ vout, int n,
                     int shift, int pass, const int *__restrict__ npass_dev, int tpb,
                     const int *__restrict__ table) {
    pdl_wait();
    pdl_trigger();
    if (npass_dev && pass >= *npass_dev) return;
    __shared__ unsigned short whist[kRadixWarps][kRadixBins];  // keys of the digit seen so far by the warp (<= 256)
    __shared__ unsigned short tile_count[kRadixBins];          // keys of the digit in the current tile (<= 2048)
    __shared__ int gbase[kRadixBins];                          // first output position of (digit, current tile)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int d = threadIdx.x; d < kRadixBins; d += kRadixBlock)
        gbase[d] = __ldg(table + (size_t)d * gridDim.x + blockIdx.x);
    for (int tile = 0; tile < tpb; ++tile) {
        const long long tile_base = ((long long)blockIdx.x * tpb + tile) * kRadixTile;
        if (tile_base >= n) break;
        for (int d = threadIdx.x; d < kRadixBins; d += kRadixBlock) {
#pragma unroll
            for (int w = 0; w < kRadixWarps; ++w) whist[w][d] = 0;
        }
        __syncthreads();
        const long long base = tile_base + warp * (kRadixTile / kRadixWarps);
        K key[kRadixRounds];
        int val[kRadixRounds];
        int rank[kRadixRounds];
#pragma unroll
        for (int r = 0; r < kRadixRounds; ++r) {
            const long long i = base + r * 32 + lane;
            const bool valid = i < n;
            key[r] = valid ? load((int)i) : (K)0;
            val[r] = valid ? (vin ? __ldg(vin + i) : (int)i) : 0;
            const int d = (int)((key[r] >> shift) & (K)(kRadixBins - 1));
            // lanes past the end get a value nobody else has: they match only themselves
            const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : (kRadixBins + lane));
            const int before = __popc(peers & ((1u << lane) - 1u));
            const int seen = whist[warp][d];
            rank[r] = seen + before;
            __syncwarp();
            if (valid && before == 0) whist[warp][d] = (unsigned short)(seen + __popc(peers));
            __syncwarp();
        }
        __syncthreads();
        // per digit: exclusive prefix over the warps (warp w's keys come after those of warps < w)
        for (int d = threadIdx.x; d < kRadixBins; d += kRadixBlock) {
            int run = 0;
#pragma unroll
            for (int w = 0; w < kRadixWarps; ++w) {
                const int t = whist[w][d];
                whist[w][d] = (unsigned short)run;
                run += t;
            }
            tile_count[d] = (unsigned short)run;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kRadixRounds; ++r) {
            const long long i = base + r * 32 + lane;
            if (i < n) {
                const int d = (int)((key[r] >> shift) & (K)(kRadixBins - 1));
                const int pos = gbase[d] + whist[warp][d] + rank[r];
                if (kout) kout[pos] = key[r];
                vout[pos] = val[r];
            }
        }
        __syncthreads();
        for (int d = threadIdx.x; d < kRadixBins; d += kRadixBlock) gbase[d] += tile_count[d];
        // (the next iteration's zeroing of whist + its barrier order these updates before they are read)
    }
}

// One pass over digit `pass` (bits [11 pass, 11 pass + 11)).  vin == NULL: the values are 0..n-1 (first pass).
// kout == NULL: keys are not written (last pass when only the permutation is wanted).
// scratch: radix_scratch_ints(n) ints.  Enqueues 3 kernels (the histogram kernel also zeroes the scan's state).
template <typename K, class Load>
inline void launch_radix_pass(Load load, const int *vin, K *kout, int *vout, int n, int pass, const int *npass_dev,
                              int *scratch, cudaStream_t st, bool pdl = false) {
    if (n <= 0) return;
    const int tpb = radix_tiles_per_block(n), blocks = radix_blocks(n);
    int *table = scratch;
    const int entries = kRadixBins * blocks;
    int *partial = scratch + radix_table_ints(n) + 1;
    if ((reinterpret_cast<uintptr_t>(partial) & 7u) != 0) ++partial;  // the scan's states are 64-bit words
    const int shift = pass * kRadixBits;
    launch_chain(pdl, radix_hist_kernel<K, Load>, blocks, kRadixBlock, 0, st, load, n, shift, pass, npass_dev, tpb, table,
                 partial, (int)scan_partial_ints(entries));
    launch_exclusive_scan_chained(table, table, entries, partial, st, pdl);
    launch_chain(pdl, radix_scatter_kernel<K, Load>, blocks, kRadixBlock, 0, st, load, vin, kout, vout, n, shift, pass,
                 npass_dev, tpb, (const int *)table);
}
constexpr int kRadixLaunchesPerPass = 3;

}  // namespace aopt
