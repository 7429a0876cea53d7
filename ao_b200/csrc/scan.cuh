// scan.cuh — exclusive prefix sum of int32 counts (used by the CSR builder and the kNN grid).
// out[i] = sum_{j<i} in[j] for i in [0, n], i.e. out has n+1 entries and out[n] is the total.
// `in` and `out` may be the same buffer.  `partial` needs div_up(n, kScanTile) + 1 ints.
#pragma once
#include "common.cuh"

namespace aopt {

constexpr int kScanBlock = 1024;
constexpr int kScanItems = 4;  // consecutive elements per thread
constexpr int kScanTile = kScanBlock * kScanItems;

inline size_t scan_partial_ints(long long n) { return (size_t)div_up(n, kScanTile) + 1; }

// Launches 3 kernels on `st`; defined in scan.cu.
void launch_exclusive_scan(const int *in, int *out, int n, int *partial, cudaStream_t st);

}  // namespace aopt
