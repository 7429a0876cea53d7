// scan.cuh — exclusive prefix sum of int32 counts (used by the CSR builder, the kNN grid and the voxel
// partition).  out[i] = sum_{j<i} in[j] for i in [0, n], i.e. out has n+1 entries and out[n] is the total.
// `in` and `out` may be the same buffer.  `partial` is scratch of scan_partial_ints(n) ints.
// Single pass (decoupled look-back): one memset + one kernel instead of three kernels.
#pragma once
#include "common.cuh"

namespace aopt {

constexpr int kScanBlock = 512;
constexpr int kScanItems = 8;  // consecutive elements per thread
constexpr int kScanTile = kScanBlock * kScanItems;

// tile states (one 64-bit word per tile) + the dynamic tile counter
inline size_t scan_partial_ints(long long n) { return 2 * ((size_t)div_up(n, kScanTile) + 2); }

// Enqueues one memset and one kernel on `st`; defined in scan.cu.
void launch_exclusive_scan(const int *in, int *out, int n, int *partial, cudaStream_t st);
// Same without the memset: the caller guarantees that the scan_partial_ints(n) ints of `partial` were zeroed by an
// earlier kernel on the stream (radix.cuh: the histogram kernel of the pass does it — one graph node less per pass).
void launch_exclusive_scan_prezeroed(const int *in, int *out, int n, int *partial, cudaStream_t st);
// Pre-zeroed state as above, launched inside a chain of programmatic dependent launches when pdl is true (common.cuh).
void launch_exclusive_scan_chained(const int *in, int *out, int n, int *partial, cudaStream_t st, bool pdl);

}  // namespace aopt
