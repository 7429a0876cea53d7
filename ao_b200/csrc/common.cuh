// common.cuh — shared device helpers for libao_pointops (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ao_pointops.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libao_pointops is written for sm_100a (B200); no other architecture is supported"
#endif

namespace aopt {

constexpr int kNumSM = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// Records the launch error for aopt_last_cuda_error() and maps it to a status; `kernels` = how many
// kernels the entry point just enqueued (summed into aopt_kernel_launches()).
int check_launch(int kernels = 1);

// Tuning switches (A/B measurements and tests): initialised from the environment variable of the same meaning on
// first use, changeable at run time through aopt_set_tuning (api.cu).  0 always means "library default".
enum Tuning { kTuneCsrImpl = 0,   // AOPT_CSR_IMPL:   1 = radix sort, 2 = count / fill / rank
              kTuneGvaBwd = 1,    // AOPT_GVA_BWD:    1 = fused kernel, 2 = two kernels
              kTuneVoxelSort = 2, // AOPT_VOXEL_SORT: 1 = own radix sort (3 passes), 2 = wide keys (6 passes)
              kTuneKnnTopk = 3,   // AOPT_KNN_TOPK:   1 = shared-memory heap (K >= 8), 0 / 2 = sorted list in registers (default)
              kTuneKnnPend = 4,   // AOPT_KNN_PEND:   1 = per-lane pending list in the GRID query kernel (K >= 8), 0 / 2 = insert in place (default)
              kTuneKnnSample = 5, // AOPT_KNN_SAMPLE: 1 = cell edge from the bounding box (no density sample), 0 / 2 = sampled r_k (default)
              kTunePdl = 6,       // AOPT_PDL:        1 = programmatic dependent launch inside the small-kernel chains, 2 = off
              kTuneL2Prefetch = 7, // AOPT_L2PF:     0 / 1 = the GVA kernels request the next work item's (k,C) peb block into L2 with a bulk prefetch (default), 2 = off
              kTuneKnnSite = 8,    // AOPT_KNN_SITE:  1 = GRID query kernel with a single insert site (knn_grid1_kernel), 0 / 2 = knn_grid_kernel (default)
              kTuneCount = 9 };
int tuning(int which);

// ---- programmatic dependent launch (PDL) for chains of small dependent kernels -----------------------------------
// A kernel launched with launch_chain(pdl = true, ...) may be scheduled while its predecessor on the stream is still
// running; it MUST call pdl_wait() before it touches global memory (the wait returns once the predecessor grid has
// completed and its writes are visible; it is a no-op for a normally launched kernel).  pdl_trigger() in the
// predecessor lets the dependent grid start launching early.  What overlaps is launch latency and ramp-up, which is most
// of what a 4-8 us kernel costs.  Tuning "pdl" (AOPT_PDL): 1 = on, 2 = off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline void launch_chain(bool pdl, void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
    if (!pdl) {
        kernel<<<grid, block, smem, st>>>(static_cast<KArgs>(args)...);
        return;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3((unsigned)block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Bulk L2 prefetch (SASS UBLKPF.L2): `bytes` (a multiple of 16) from a 16-byte aligned global address are requested into L2
// by the copy engine — no LSU traffic, no destination, no completion to wait for.  The address is warp-uniform in hardware:
// ptxas wraps a per-lane address in a loop over the active lanes, so call it from a few lanes per warp only.
__device__ __forceinline__ void l2_prefetch_bulk(const void *p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

inline cudaStream_t as_stream(aopt_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// Grid for a grid-stride loop over `work` items: enough CTAs to cover the work once, capped at
// `ctas_per_sm` resident CTAs on each of the 148 SMs (so big problems run as full waves).
inline int stride_grid(long long work, int block, int ctas_per_sm) {
    long long need = (work + block - 1) / block;
    long long cap = (long long)kNumSM * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- 128-bit global memory access with cache hints -------------------------------------------
// Streaming (touched once) data bypasses L1 so that the gathered rows keep the cache.
__device__ __forceinline__ float4 ldg_stream4(const float *p) {
    float4 v;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream4(float *p, const float4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float ldg_stream1(const float *p) {
    float v;
    asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream1(float *p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
// Ordered variants for software-batched loads: `asm volatile` statements keep their program order, so a
// run of these followed by issue_fence() reaches ptxas as "all loads, then the consumers" — the
// compiler otherwise sinks each load next to its use (fewer live registers, but only ~3 requests in
// flight per thread, and these kernels are bound by DRAM latency, not by issue slots).
__device__ __forceinline__ float4 ldg_stream4_ordered(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldg_gather4_ordered(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void issue_fence() { asm volatile("" ::: "memory"); }

// ---- per-thread asynchronous staging (cp.async → LDGSTS) ----------------------------------------
// A thread copies 16-byte pieces straight from global to its OWN shared-memory slots and reads them
// back after cp.async.wait_group: no registers are tied up while the requests are in flight and no
// barrier is needed (nobody else touches the slots).  Slot u of thread t lives at stage[u*blockDim+t]
// (consecutive threads → consecutive 16-byte words: conflict-free for the copy and for the read).
__device__ __forceinline__ void cp_async16_stream(float4 *smem_dst, const float *gsrc) {  // L2 only
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_gather(float4 *smem_dst, const float *gsrc) {  // keep in L1
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Gathered rows: read-only path, default caching (re-used across neighbouring queries).
__device__ __forceinline__ float4 ldg_gather4(const float *p) {
    return __ldg(reinterpret_cast<const float4 *>(p));
}

// A VEC-wide channel chunk (VEC = 4: one 128-bit access; VEC = 1: scalar fallback for widths that
// are not a multiple of 4 or for mis-aligned strides such as the (m,k,3+c) concatenated layout).
template <int VEC>
struct Chunk;
template <>
struct Chunk<4> {
    float4 v;
    __device__ __forceinline__ static Chunk zero() { return {make_float4(0.f, 0.f, 0.f, 0.f)}; }
    __device__ __forceinline__ static Chunk gather(const float *p) { return {ldg_gather4(p)}; }
    __device__ __forceinline__ static Chunk stream(const float *p) { return {ldg_stream4(p)}; }
    __device__ __forceinline__ void store_stream(float *p) const { stg_stream4(p, v); }
    __device__ __forceinline__ void store(float *p) const { *reinterpret_cast<float4 *>(p) = v; }
    __device__ __forceinline__ void add(const Chunk &o) { v.x += o.v.x; v.y += o.v.y; v.z += o.v.z; v.w += o.v.w; }
    __device__ __forceinline__ void sub(const Chunk &o) { v.x -= o.v.x; v.y -= o.v.y; v.z -= o.v.z; v.w -= o.v.w; }
    __device__ __forceinline__ void scale(float s) { v.x *= s; v.y *= s; v.z *= s; v.w *= s; }
    __device__ __forceinline__ void fma(const Chunk &a, float s) {
        v.x = fmaf(a.v.x, s, v.x); v.y = fmaf(a.v.y, s, v.y); v.z = fmaf(a.v.z, s, v.z); v.w = fmaf(a.v.w, s, v.w);
    }
};
template <>
struct Chunk<1> {
    float v;
    __device__ __forceinline__ static Chunk zero() { return {0.f}; }
    __device__ __forceinline__ static Chunk gather(const float *p) { return {__ldg(p)}; }
    __device__ __forceinline__ static Chunk stream(const float *p) { return {ldg_stream1(p)}; }
    __device__ __forceinline__ void store_stream(float *p) const { stg_stream1(p, v); }
    __device__ __forceinline__ void store(float *p) const { *p = v; }
    __device__ __forceinline__ void add(const Chunk &o) { v += o.v; }
    __device__ __forceinline__ void sub(const Chunk &o) { v -= o.v; }
    __device__ __forceinline__ void scale(float s) { v *= s; }
    __device__ __forceinline__ void fma(const Chunk &a, float s) { v = fmaf(a.v, s, v); }
};

// Fast division of a 64-bit work index by a small runtime constant (channel chunks per row):
// done once per item; the kernels are HBM-bound so a hardware divide is far below the issue budget.
struct RowCol {
    long long row;
    int col;
};
__device__ __forceinline__ RowCol split(long long t, int cols) {
    RowCol rc;
    rc.row = t / cols;
    rc.col = (int)(t - rc.row * cols);
    return rc;
}

// ---- "thread keeps its column" grid-stride walk over (rows x cols) items -----------------------
// t -> (t / cols, t % cols) costs a 64-bit division per item; for the kernels that do one small gather per item
// (interpolation forward, pool backward) that is more instructions than the work itself.  When gridDim.x * block is
// a multiple of cols, a thread's column never changes and only its row advances by a constant:
//   ColWalk w = col_walk(cols, BLOCK);  for (long long row = w.row; row < n_rows; row += w.row_step) { ... w.col ... }
inline int gcd_i(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }
inline int col_grid(long long n_rows, int cols, int block, int ctas_per_sm) {
    const int unit = cols / gcd_i(cols, block);   // CTAs per whole number of rows
    long long need = (n_rows * cols + block - 1) / block;
    long long cap = (long long)kNumSM * ctas_per_sm;
    // below the cap: round UP (every row covered by the first pass); at the cap: round DOWN — one CTA more than
    // a full wave runs its whole row walk alone after everybody else has finished (pool_backward: 39 -> 51 us)
    long long grid = ((need + unit - 1) / unit) * unit;
    if (grid > cap) grid = (cap / unit) * unit;
    if (grid < unit) grid = unit;
    return (int)grid;
}
struct ColWalk {
    long long row, row_step;
    int col;
};
__device__ __forceinline__ ColWalk col_walk(int cols, int block) {
    const long long t0 = (long long)blockIdx.x * block + threadIdx.x;
    ColWalk w;
    w.row = t0 / cols;
    w.col = (int)(t0 - w.row * cols);
    w.row_step = ((long long)gridDim.x * block) / cols;
    return w;
}

}  // namespace aopt
