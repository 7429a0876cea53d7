// csr.cu — transpose of the neighbour graph: for every source point j the list of flat
// (query, slot) positions p = m*k+s that reference it, in ascending p.
//
// This is what makes every backward of the path atomic-free and run-to-run deterministic.  The
// reference scatters gradients with float atomicAdd
// (libs/pointops/src/grouping/grouping_cuda_kernel.cu:24, interpolation_cuda_kernel.cu:31,
// aggregation_cuda_kernel.cu:35-37), which is contended on hub points and non-deterministic.
// One CSR is built per kNN result and shared by all blocks of a BlockSequence
// (pointcept/models/point_transformer_v2/point_transformer_v2m2_base.py:223-225 reuses one
// reference_index for every block).
//
// Steps (all on the caller's stream, integer work only):
//   1. count[j]   += 1 per entry            (int atomics: the counts are order-independent)
//   2. rowptr      = exclusive scan(count)  (scan.cu, single pass)
//   3. tmp[rowptr[j] + slot] = p            (slot = what the atomic of step 1 returned: arbitrary order inside a row)
//   4. rank every entry inside its row by the value of p (all-pairs count inside the row, rows are
//      ~k long) and write perm[rowptr[j] + rank] = p   → ascending p, deterministic.
// Workspace: count (n_src+1 ints) + tmp and slot (n_entries ints each) + scan partials.
//
// Hub-heavy graphs take a different route to the same arrays (build_sorted below): a stable radix sort of the flat
// positions by source index (radix.cuh).  Stability IS the ascending-p order inside a row, there are no atomics, and a
// hub row costs what any other entries cost (the rank of step 4 is quadratic in the in-degree).
#include <stdlib.h>

#include "common.cuh"
#include "radix.cuh"
#include "scan.cuh"

namespace aopt {

constexpr int kBlock = 256;

// Sort key of entry p: its source row, or n_src for entries that reference nothing (-1 padding, out of range):
// those sort to the end and rowptr[n_src] = number of kept entries.
struct IdxKeys {
    const int *idx;
    int n_src, negative_mode;
    __device__ __forceinline__ unsigned operator()(int p) const {
        int j = __ldg(idx + p);
        if (j < 0 && negative_mode == 1) j += n_src;
        return (j >= 0 && j < n_src) ? (unsigned)j : (unsigned)n_src;
    }
};

// rowptr[j] = first sorted position whose key is >= j (j = 0..n_src): a binary search per row, so empty rows —
// however many in a run — cost nothing extra.
__global__ void __launch_bounds__(kBlock)
csr_rowptr_kernel(int n_src, int n_entries, const unsigned *__restrict__ sorted_keys, int *__restrict__ rowptr) {
    const int j = blockIdx.x * kBlock + threadIdx.x;
    if (j > n_src) return;
    int lo = 0, hi = n_entries;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(sorted_keys + mid) < (unsigned)j) lo = mid + 1;
        else hi = mid;
    }
    rowptr[j] = lo;
}

__device__ __forceinline__ int wrap_index(int j, int n_src, int negative_mode) {
    return (j < 0 && negative_mode == 1) ? j + n_src : j;
}

// count[j] += 1 per entry; the value the atomic returns is the entry's (arbitrary) slot inside its row, kept so
// that the fill needs no second round of atomics (it was 70 us of the 165 us level-0 build).
__global__ void __launch_bounds__(kBlock)
csr_count_kernel(long long n_entries, int n_src, int negative_mode, const int *__restrict__ idx,
                 int *__restrict__ count, int *__restrict__ slot, int *__restrict__ scan_state, int scan_state_ints) {
    pdl_trigger();
    // the look-back states of the scan that follows (launch_exclusive_scan_chained): zeroed here, one memset node less
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < scan_state_ints; i += gridDim.x * kBlock) scan_state[i] = 0;
    const long long step = (long long)gridDim.x * kBlock;
    for (long long p = (long long)blockIdx.x * kBlock + threadIdx.x; p < n_entries; p += step) {
        int j = wrap_index(__ldg(idx + p), n_src, negative_mode);
        if (j >= 0 && j < n_src) slot[p] = atomicAdd(count + j, 1);
    }
}

__global__ void __launch_bounds__(kBlock)
csr_fill_kernel(long long n_entries, int n_src, int negative_mode, const int *__restrict__ idx,
                const int *__restrict__ rowptr, const int *__restrict__ slot, int *__restrict__ tmp) {
    pdl_wait();
    pdl_trigger();
    const long long step = (long long)gridDim.x * kBlock;
    for (long long p = (long long)blockIdx.x * kBlock + threadIdx.x; p < n_entries; p += step) {
        int j = wrap_index(__ldg(idx + p), n_src, negative_mode);
        if (j >= 0 && j < n_src) tmp[__ldg(rowptr + j) + slot[p]] = (int)p;
    }
}

// One thread per filled entry e: rank of tmp[e] among the entries of its row.
__global__ void __launch_bounds__(kBlock)
csr_rank_kernel(int n_src, int negative_mode, const int *__restrict__ idx,
                const int *__restrict__ rowptr, const int *__restrict__ tmp, int *__restrict__ perm) {
    pdl_wait();
    const int total = __ldg(rowptr + n_src);
    const int step = gridDim.x * kBlock;
    for (int e = blockIdx.x * kBlock + threadIdx.x; e < total; e += step) {
        int p = tmp[e];
        int j = wrap_index(__ldg(idx + p), n_src, negative_mode);
        int b = __ldg(rowptr + j), end = __ldg(rowptr + j + 1);
        int rank = 0;
        for (int q = b; q < end; ++q) rank += (tmp[q] < p) ? 1 : 0;
        perm[b + rank] = p;
    }
}

}  // namespace aopt

using namespace aopt;

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Which builder (tuning "csr_impl" / AOPT_CSR_IMPL: 1 = sort, 2 = count).  Measured on a B200 (profiles/r02a_kernel_bench.txt,
// kNN graphs, in-degree ~16): count / fill / rank 184 us vs 283 us for the two-pass sort at 5.1 M entries, 54 vs 113 us at
// 0.8 M — with rows this short the all-pairs rank is cheap and the sort pays two full passes over the entries.  The rank
// is quadratic in the row length, though, so graphs whose AVERAGE in-degree is large (a coarse level interpolated to a
// fine one, a generic `cluster` map) take the sort, whose cost does not depend on the degree distribution.
constexpr long long kSortMinAvgDegree = 64;
static bool use_sort(int n_src, long long n_entries) {
    const int m = tuning(kTuneCsrImpl);
    return m == 1 || (m == 0 && n_entries >= kSortMinAvgDegree * (long long)(n_src > 0 ? n_src : 1) && n_entries >= 4096);
}

static size_t count_workspace(int n_src, int64_t n_entries) {
    return align256(((size_t)n_src + 1) * 4) + 2 * align256((size_t)n_entries * 4) +
           align256(scan_partial_ints(n_src) * 4);
}
static size_t sort_workspace(int64_t n_entries) {   // two key buffers, two value buffers, the pass scratch
    return 4 * align256((size_t)n_entries * 4) + align256(radix_scratch_ints(n_entries) * 4 + 16);
}

extern "C" size_t aopt_csr_workspace_bytes(int n_src, int64_t n_entries) {
    if (n_src < 0 || n_entries < 0) return 0;
    const size_t a = count_workspace(n_src, n_entries), b = sort_workspace(n_entries);
    return a > b ? a : b;
}

// Stable radix sort of the entries by source row; 1-3 passes of 11 bits cover n_src (the padding key).
static int build_sorted(int n_src, int n_entries, const int *idx, int negative_mode, int *rowptr, int *perm,
                        char *ws, cudaStream_t st) {
    unsigned *k0 = reinterpret_cast<unsigned *>(ws); ws += align256((size_t)n_entries * 4);
    unsigned *k1 = reinterpret_cast<unsigned *>(ws); ws += align256((size_t)n_entries * 4);
    int *v0 = reinterpret_cast<int *>(ws); ws += align256((size_t)n_entries * 4);
    int *v1 = reinterpret_cast<int *>(ws); ws += align256((size_t)n_entries * 4);
    int *scratch = reinterpret_cast<int *>(ws);
    int bits = 1;
    while (bits < 32 && (((unsigned)n_src) >> bits) != 0) ++bits;
    const int npass = (bits + kRadixBits - 1) / kRadixBits;
    const IdxKeys first{idx, n_src, negative_mode};
    unsigned *kin = nullptr, *kbuf[2] = {k0, k1};
    int *vin = nullptr, *vbuf[2] = {v0, v1};
    for (int pass = 0; pass < npass; ++pass) {
        const bool last = pass == npass - 1;
        unsigned *kout = kbuf[pass & 1];
        int *vout = last ? perm : vbuf[pass & 1];
        if (pass == 0) launch_radix_pass<unsigned>(first, nullptr, kout, vout, n_entries, pass, nullptr, scratch, st);
        else launch_radix_pass<unsigned>(PtrKeys<unsigned>{kin}, vin, kout, vout, n_entries, pass, nullptr, scratch, st);
        kin = kout; vin = vout;
    }
    csr_rowptr_kernel<<<div_up((long long)n_src + 1, kBlock), kBlock, 0, st>>>(n_src, n_entries, kin, rowptr);
    return check_launch(npass * kRadixLaunchesPerPass + 1);
}

extern "C" int aopt_csr_build(int n_src, int64_t n_entries, const int *idx, int negative_mode,
                              int *rowptr, int *perm, void *workspace, size_t workspace_bytes,
                              aopt_stream_t stream) {
    if (n_src < 0 || n_entries < 0 || n_entries > 0x7fffffffLL || (negative_mode != 0 && negative_mode != 1))
        return AOPT_ERR_INVALID_ARGUMENT;
    if (!rowptr) return AOPT_ERR_INVALID_ARGUMENT;
    cudaStream_t st = as_stream(stream);
    if (n_src == 0) {
        cudaMemsetAsync(rowptr, 0, sizeof(int), st);
        return check_launch(0);
    }
    if (n_entries > 0 && (!idx || !perm)) return AOPT_ERR_INVALID_ARGUMENT;
    if (!workspace || workspace_bytes < aopt_csr_workspace_bytes(n_src, n_entries)) return AOPT_ERR_WORKSPACE;
    char *ws = static_cast<char *>(workspace);
    if (n_entries > 0 && use_sort(n_src, n_entries))
        return build_sorted(n_src, (int)n_entries, idx, negative_mode, rowptr, perm, ws, st);
    int *count = reinterpret_cast<int *>(ws);
    ws += align256(((size_t)n_src + 1) * 4);
    int *tmp = reinterpret_cast<int *>(ws);
    ws += align256((size_t)n_entries * 4);
    int *slot = reinterpret_cast<int *>(ws);
    ws += align256((size_t)n_entries * 4);
    int *partial = reinterpret_cast<int *>(ws);

    cudaMemsetAsync(count, 0, ((size_t)n_src + 1) * 4, st);
    if (n_entries == 0) {
        launch_exclusive_scan(count, rowptr, n_src, partial, st);
        return check_launch(1);
    }
    // count -> scan -> fill -> rank as a chain of programmatic dependent launches (tuning "pdl")
    const bool pdl = tuning(kTunePdl) != 2;
    const int grid = stride_grid(n_entries, kBlock, 8);
    csr_count_kernel<<<grid, kBlock, 0, st>>>(n_entries, n_src, negative_mode, idx, count, slot, partial,
                                              (int)scan_partial_ints(n_src));
    launch_exclusive_scan_chained(count, rowptr, n_src, partial, st, pdl);   // rowptr[n_src] = number of kept entries
    launch_chain(pdl, csr_fill_kernel, grid, kBlock, 0, st, n_entries, n_src, negative_mode, idx, (const int *)rowptr,
                 (const int *)slot, tmp);
    launch_chain(pdl, csr_rank_kernel, grid, kBlock, 0, st, n_src, negative_mode, idx, (const int *)rowptr, (const int *)tmp, perm);
    return check_launch(n_entries > 0 ? 4 : 1);  // count, scan, fill, rank
}
