// bbox.cuh — per-scene bounding boxes of the offset-encoded batch, computed by the whole chip
// (the one-block-per-scene form took 57 us for 4 x 80k points; this one is ~5 us).
// Floats are accumulated with integer atomics on an order-preserving encoding.
#pragma once
#include "common.cuh"
#include "knn_common.cuh"

namespace aopt {

// Monotone float → unsigned map: a < b  ⇔  enc(a) < enc(b)  (for non-NaN values).
__device__ __forceinline__ unsigned bbox_encode(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float bbox_decode(unsigned e) {
    return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}
constexpr unsigned kBboxEmptyLo = 0xffffffffu;  // "no point seen" markers (memset patterns 0xff / 0x00)
constexpr unsigned kBboxEmptyHi = 0u;

// lo / hi: (b,3) unsigned encodings.  hi may be NULL (minimum only).  Enqueues the memsets too.
// init = false: the caller has already set lo to 0xffffffff and hi to 0 on this stream (the grid kNN folds
// that into its sample kernel: three fewer memset nodes per search).
void launch_scene_bbox(int n, int b, const float *xyz, const int *offset, unsigned *lo, unsigned *hi,
                       cudaStream_t st, bool init = true, bool pdl = false);

}  // namespace aopt
