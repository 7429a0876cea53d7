// knn_grid.cu — exact kNN with spatial pruning (the GRID method of aopt_knn_query).
//
// The reference (and the TILE method) scan every candidate of the scene for every query:
// 80k x 80k pairs per S3DIS room, ALU-bound.  Here each scene gets a uniform grid
// sized from the data (cell edge ~ the k-th neighbour distance measured on a sample), candidates
// are counting-sorted by cell, and each query visits the 3x3x3 block around its cell, then
// successive shells, until the k-th best distance is provably smaller than the distance to
// anything outside the visited block.  The result is the same set, in the same (d2, idx) order,
// as the exhaustive scan: distances use the same dist2_ref() and candidates are ranked
// lexicographically, so the traversal order does not matter.
//
// Pipeline (all on the caller's stream, no host synchronisation, no allocation):
//   1 knn_sample_kernel   r_k^2 estimate at 64 sample points per scene (one CTA per sample, thinned
//                         candidate set, register top-k + block-wide pops)
//   2 scene_bbox_kernel + grid_setup_kernel   per-scene bounding box (bbox.cu), cell edge h, dims → GridDesc
//   3 grid_count_kernel   cell of every candidate + its slot inside the cell (int atomics)
//   4 exclusive scan      cell counts → cell starts                         (scan.cu)
//   5 grid_fill_kernel    candidates → cell order as float4 (x, y, z, original index)
//   6 knn_grid_kernel<K>  one thread per query (self-queries run in cell order so that a warp
//                         shares its candidate runs through L1), register top-k, LEX ranking
//
// Exactness of the stopping rule.  After visiting the block of cells [c-r, c+r]^3 (clipped to the
// grid) every unvisited candidate lies beyond one of the block's inner faces.  bound = the smallest
// distance from the query to such a face.  Cell indices are computed in fp32, so a candidate may
// sit up to ~1e-6·dim cells on the wrong side of a face; the rule therefore stops only when
// kth_d2 < (bound - margin)^2 with margin = h·(1e-3 + 1e-6·max_dim) — three orders of magnitude
// more than the rounding it has to absorb.  Being conservative costs an extra shell, never
// correctness.  The strict '<' also rules out an unvisited candidate tying with the k-th best.
#include <stdlib.h>

#include "common.cuh"
#include "knn_common.cuh"
#include "scan.cuh"
#include "bbox.cuh"

namespace aopt {

constexpr int kSamples = 64;        // sample points per scene for the density estimate
// grid capacity: cells per candidate (+ 64 per scene).  Indoor scans are surfaces: most cells of the volume are empty, so the
// capacity — not the sampled k-th neighbour distance — usually fixes the cell edge.  AOPT_KNN_CELLS_PER_POINT overrides (sweeps).
static int cells_per_point() {
    static const int v = [] {
        int d = 4;
        if (const char *e = getenv("AOPT_KNN_CELLS_PER_POINT")) { int x = atoi(e); if (x >= 1 && x <= 64) d = x; }
        return d;
    }();
    return v;
}
constexpr int kCellsPerScene = 64;
// Beyond this shell radius a query falls back to a full scan of its scene.  8 was too eager: in the outdoor scans
// (SemanticKITTI-shaped, configs[4]) the far-range points need many shells of the cell size set by the dense near range,
// and the k = 32 search spent 8.4 ms in 80k-candidate scans (0.22 ms at k = 16, profiles/r02a).  Walking the shells up
// to radius r costs ~ (8/3) r^3 cell-range lookups: 21k at r = 20, still far below a scan of the whole scene.
constexpr int kMaxRing = 20;
constexpr int kQueryBlock = 128;

struct GridDesc {
    float ox, oy, oz;  // bbox minimum of the scene's candidates
    float h, inv_h;    // cell edge
    int nx, ny, nz;
    int cell_base;     // first cell of the scene in the global cell array
    int start, end;    // candidate range of the scene
    float margin;
    int pad[4];
};
static_assert(sizeof(GridDesc) == 64, "GridDesc is 64 bytes");

__device__ __forceinline__ int cell_coord(float p, float o, float inv_h, int dim) {
    int c = (int)floorf((p - o) * inv_h);
    return min(max(c, 0), dim - 1);
}

// ---- 1. density sample ---------------------------------------------------------------------------
// One CTA per sample point.  Big scenes are thinned by an index stride (the cell edge only steers
// speed, never correctness): the (k-1)/stride-th nearest neighbour in the thinned set estimates the
// k-th nearest in the full set.  Every thread keeps a register top-K of its share of the candidates
// (loads for four candidates are issued together), then the CTA pops the block-wide minimum kp times.
constexpr int kSampleBlock = 512;   // 128 threads left every sample a ~30 us serial chain (profiles/r01i); same result at any width
constexpr int kSampleTarget = 8192;  // candidates scanned per sample after thinning

template <int K>
__global__ void __launch_bounds__(kSampleBlock)
knn_sample_kernel(int b, int nsample, const float *__restrict__ xyz, const int *__restrict__ offset,
                  float *__restrict__ samples, int *__restrict__ cells, long long n_cells,
                  unsigned *__restrict__ bb_lo, unsigned *__restrict__ bb_hi) {
    __shared__ float wmin[kSampleBlock / 32];
    pdl_trigger();
    // first kernel of the search on the stream: it also clears the cell counters and resets the bounding boxes
    // (work the following kernels need done; three memset nodes less per search)
    for (long long i = (long long)blockIdx.x * kSampleBlock + threadIdx.x; i < n_cells; i += (long long)gridDim.x * kSampleBlock)
        cells[i] = 0;
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < 3 * b; i += kSampleBlock) { bb_lo[i] = kBboxEmptyLo; bb_hi[i] = kBboxEmptyHi; }
    const int sc = blockIdx.x / kSamples, s = blockIdx.x - sc * kSamples;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int start = sc == 0 ? 0 : __ldg(offset + sc - 1), end = __ldg(offset + sc);
    const int ns = end - start;
    if (ns <= 0) {
        if (threadIdx.x == 0) samples[blockIdx.x] = 1e10f;
        return;
    }
    const int stride = max(1, min(ns / kSampleTarget, max(1, (nsample - 1) / 2)));
    const int kp = max(1, min(K, (max(nsample - 1, 1) + stride - 1) / stride));  // neighbours excluding the sample itself
    const int qi = start + (int)(((long long)(2 * s + 1) * ns) / (2 * kSamples));
    const float qx = __ldg(xyz + (size_t)qi * 3), qy = __ldg(xyz + (size_t)qi * 3 + 1),
                qz = __ldg(xyz + (size_t)qi * 3 + 2);
    TopK<K, false> top;
    top.init();
    const int cnt = (ns + stride - 1) / stride;  // thinned candidates: start + stride * t
    int t = threadIdx.x;
    for (; t + 3 * kSampleBlock < cnt; t += 4 * kSampleBlock) {
        float d[4];
        int id[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            id[u] = start + (t + u * kSampleBlock) * stride;
            const float *p = xyz + (size_t)id[u] * 3;
            d[u] = dist2_ref(qx, qy, qz, __ldg(p), __ldg(p + 1), __ldg(p + 2));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (id[u] != qi) top.offer(d[u], id[u]);
    }
    for (; t < cnt; t += kSampleBlock) {
        const int id = start + t * stride;
        const float *p = xyz + (size_t)id * 3;
        if (id != qi) top.offer(dist2_ref(qx, qy, qz, __ldg(p), __ldg(p + 1), __ldg(p + 2)), id);
    }
    float result = 1e10f;
    for (int r = 0; r < kp; ++r) {
        const float head = top.d[0];
        float mn = head;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
        const unsigned who = __ballot_sync(0xffffffffu, head == mn);
        if (lane == 0) wmin[warp] = mn;
        __syncthreads();
        float bm = wmin[0];
        int bw = 0;
#pragma unroll
        for (int w = 1; w < kSampleBlock / 32; ++w)
            if (wmin[w] < bm) { bm = wmin[w]; bw = w; }
        if (warp == bw && lane == __ffs(who) - 1) {
#pragma unroll
            for (int i = 0; i + 1 < K; ++i) top.d[i] = top.d[i + 1];
            top.d[K - 1] = 1e10f;
        }
        result = bm;
        __syncthreads();
    }
    if (threadIdx.x == 0) samples[blockIdx.x] = result;
}

// ---- 1b. no density sample: clear only ---------------------------------------------------------------
// Experiment kept behind tuning "knn_sample" = 1 (AOPT_KNN_SAMPLE=bbox): cell edge from the bounding box alone,
// h = cbrt(volume / capacity) per scene, this kernel doing only the sample kernel's housekeeping (21 us -> 3 us per
// search).  Measured (profiles/r02d_kernel_bench*.txt): the query pays more than the build saves — level-0 self search
// 426 us vs 328 us with the sampled edge, level 2 185 vs 130 us, the step 8.06 vs 7.91 ms — so the sample stays.
__global__ void __launch_bounds__(256)
knn_clear_kernel(int b, int *__restrict__ cells, long long n_cells, unsigned *__restrict__ bb_lo, unsigned *__restrict__ bb_hi) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n_cells; i += (long long)gridDim.x * 256) cells[i] = 0;
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < 3 * b; i += 256) { bb_lo[i] = kBboxEmptyLo; bb_hi[i] = kBboxEmptyHi; }
}

// ---- 2. per-scene grid descriptor ------------------------------------------------------------------
// One thread per scene: bounding box from the encoded min/max (bbox.cu), cell edge from the samples.
// One WARP per scene (the 64 sample loads are issued together and reduced with shuffles; one thread per scene
// walked them one after the other: 10 us per launch for 4 scenes); lane 0 derives the descriptor.
__global__ void __launch_bounds__(128)
grid_setup_kernel(int b, int n, const int *__restrict__ offset, const unsigned *__restrict__ bb_lo,
                  const unsigned *__restrict__ bb_hi, const float *__restrict__ samples, float cell_scale,
                  int kCellsPerPoint, GridDesc *__restrict__ desc) {
    pdl_wait();
    pdl_trigger();
    const int sc = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (sc >= b) return;
    float sum = 0.f;
    int cnt = 0;
#pragma unroll
    for (int s = lane; s < kSamples; s += 32) {
        const float v = samples ? __ldg(samples + sc * kSamples + s) : 1e10f;
        if (v < 1e9f) { sum += sqrtf(v); ++cnt; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    }
    if (lane != 0) return;
    int start = sc == 0 ? 0 : __ldg(offset + sc - 1), end = __ldg(offset + sc);
    start = max(start, 0);
    end = min(end, n);
    const int ns = max(end - start, 0);
    float lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
        const unsigned el = bb_lo[sc * 3 + a], eh = bb_hi[sc * 3 + a];
        const bool empty = ns == 0 || el == kBboxEmptyLo;
        lo[a] = empty ? 0.f : bbox_decode(el);
        hi[a] = empty ? 0.f : bbox_decode(eh);
    }
    GridDesc g;
    float ext[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
    float max_ext = fmaxf(ext[0], fmaxf(ext[1], ext[2]));
    float h = cnt > 0 ? cell_scale * sum / (float)cnt : 0.f;
    if (!samples) {
        // capacity-driven edge: the cube root of (bounding-box volume / cells allowed); flat boxes (a wall, a ground
        // plane) use the two or one extents that are not degenerate
        const long long cap0 = (long long)kCellsPerPoint * ns + kCellsPerScene;
        float dims[3] = {ext[0], ext[1], ext[2]};
        const float tiny = max_ext * 1e-3f;
        float vol = 1.f;
        int nd = 0;
        for (int a = 0; a < 3; ++a) if (dims[a] > tiny) { vol *= dims[a]; ++nd; }
        h = nd == 3 ? cbrtf(vol / (float)cap0) : nd == 2 ? sqrtf(vol / (float)cap0) : nd == 1 ? vol / (float)cap0 : 0.f;
    }
    if (!(h > max_ext * (1.f / 2048.f))) h = max_ext * (1.f / 2048.f);  // also catches NaN / 0
    if (!(h > 1e-12f)) h = 1.f;                                         // all points coincide
    const long long cap = (long long)kCellsPerPoint * ns + kCellsPerScene;
    int nx = 1, ny = 1, nz = 1;
    for (int it = 0; it < 200; ++it) {
        nx = (int)(ext[0] / h) + 1; ny = (int)(ext[1] / h) + 1; nz = (int)(ext[2] / h) + 1;
        if ((long long)nx * ny * nz <= cap) break;
        h *= 1.25f;
    }
    g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2];
    g.h = h; g.inv_h = 1.f / h;
    g.nx = nx; g.ny = ny; g.nz = nz;
    g.cell_base = kCellsPerPoint * start + kCellsPerScene * sc;
    g.start = start; g.end = end;
    g.margin = h * (1e-3f + 1e-6f * (float)max(nx, max(ny, nz)));
    g.pad[0] = g.pad[1] = g.pad[2] = g.pad[3] = 0;
    desc[sc] = g;
}

// ---- 3. count ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
grid_count_kernel(int n, int b, const float *__restrict__ xyz, const int *__restrict__ offset,
                  const GridDesc *__restrict__ desc, int *__restrict__ cells,
                  int *__restrict__ point_cell, int *__restrict__ point_slot, int *__restrict__ scan_state, int scan_state_ints) {
    pdl_wait();
    pdl_trigger();
    for (int t = blockIdx.x * 256 + threadIdx.x; t < scan_state_ints; t += gridDim.x * 256) scan_state[t] = 0;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int sc = find_segment(i, offset, b);
    if (sc >= b) { point_cell[i] = -1; return; }
    const GridDesc g = desc[sc];
    const int cx = cell_coord(__ldg(xyz + (size_t)i * 3 + 0), g.ox, g.inv_h, g.nx);
    const int cy = cell_coord(__ldg(xyz + (size_t)i * 3 + 1), g.oy, g.inv_h, g.ny);
    const int cz = cell_coord(__ldg(xyz + (size_t)i * 3 + 2), g.oz, g.inv_h, g.nz);
    const int cell = g.cell_base + (cz * g.ny + cy) * g.nx + cx;
    point_cell[i] = cell;
    point_slot[i] = atomicAdd(cells + cell, 1);
}

// ---- 5. fill ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
grid_fill_kernel(int n, const float *__restrict__ xyz, const int *__restrict__ cells,
                 const int *__restrict__ point_cell, const int *__restrict__ point_slot,
                 float4 *__restrict__ sorted) {
    pdl_wait();
    pdl_trigger();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int cell = point_cell[i];
    if (cell < 0) return;
    const int pos = cells[cell] + point_slot[i];
    sorted[pos] = make_float4(__ldg(xyz + (size_t)i * 3), __ldg(xyz + (size_t)i * 3 + 1),
                              __ldg(xyz + (size_t)i * 3 + 2), __int_as_float(i));
}

// ---- 6. query ------------------------------------------------------------------------------------
// Two interchangeable top-k containers (same results: both rank candidates by the 64-bit (d2 bits : idx) key):
//   TopK<K, true>  (knn_common.cuh)  sorted list in registers; an insert is K compares + 2K selects on 64-bit keys
//                                    = 96 ALU-pipe instructions at K = 16, executed by the whole warp whenever ANY of its
//                                    32 queries accepts a candidate — which is nearly every candidate (an acceptance
//                                    probability of 0.1-0.3 per lane): ncu r01q: ALU pipe 81 %, 14.5 of 32 lanes active.
//   HeapK<K>       (below)           binary max-heap, one shared-memory column per thread (slot s of thread t at
//                                    [s * kQueryBlock + t]: every lane hits its own bank whatever slot it is at), root
//                                    mirrored in registers.  An insert replaces the root and sifts down: <= log2 K levels
//                                    of (4 LDS, 4 integer compares, 3 selects, 2 STS) — about a third of the ALU-pipe work,
//                                    the rest moved to the otherwise idle LSU pipe; no K x 2 registers for the list, so no
//                                    spills and more warps to hide the chain.  A heap sort at the end writes the ascending
//                                    list the API promises.  This is the reference's own data structure
//                                    (knn_query_cuda_kernel.cu:15-43) moved from local memory into conflict-free shared memory.
template <int K>
struct HeapK {
    unsigned *hd;  // column of this thread: element s at hd[s * kQueryBlock]
    int *hi;
    unsigned rd;   // root = current k-th best, mirrored
    int ri;

    static __device__ __forceinline__ unsigned long long key(unsigned dv, int iv) {
        return ((unsigned long long)dv << 32) | (unsigned)iv;
    }
    __device__ __forceinline__ void init(unsigned *sd, int *si) {
        hd = sd + threadIdx.x;
        hi = si + threadIdx.x;
        reset();
    }
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int s = 0; s < K; ++s) { hd[s * kQueryBlock] = __float_as_uint(1e10f); hi[s * kQueryBlock] = -1; }
        rd = __float_as_uint(1e10f);
        ri = -1;
    }
    __device__ __forceinline__ float worst() const { return __uint_as_float(rd); }

    // (nd, ni) enters at the root and sinks below every larger child; `limit` = heap size
    __device__ __forceinline__ void sift(unsigned nd, int ni, int limit) {
        const unsigned long long nk = key(nd, ni);
        int pos = 0;
        for (;;) {
            const int l = 2 * pos + 1;
            if (l >= limit) break;
            const unsigned dl = hd[l * kQueryBlock];
            const int il = hi[l * kQueryBlock];
            unsigned dr = 0u;
            int ir = 0;
            if (l + 1 < limit) { dr = hd[(l + 1) * kQueryBlock]; ir = hi[(l + 1) * kQueryBlock]; }
            const bool rbig = key(dr, ir) > key(dl, il);
            const unsigned dc = rbig ? dr : dl;
            const int ic = rbig ? ir : il;
            if (!(key(dc, ic) > nk)) break;
            hd[pos * kQueryBlock] = dc;
            hi[pos * kQueryBlock] = ic;
            if (pos == 0) { rd = dc; ri = ic; }
            pos = l + (rbig ? 1 : 0);
        }
        hd[pos * kQueryBlock] = nd;
        hi[pos * kQueryBlock] = ni;
        if (pos == 0) { rd = nd; ri = ni; }
    }
    __device__ __forceinline__ void offer(float d2, int i) {
        // same acceptance rule as TopK<K, true>::offer
        const unsigned du = __float_as_uint(d2);
        if (d2 < 1e10f && key(du, i) < key(rd, ri)) sift(du, i, K);
    }
    // heap sort in place (ascending), then the first nsample entries go out
    __device__ __forceinline__ void store(int *idx_row, float *d2_row, int nsample, bool root) {
        for (int end = K - 1; end >= 1; --end) {
            const unsigned nd = hd[end * kQueryBlock];
            const int ni = hi[end * kQueryBlock];
            hd[end * kQueryBlock] = rd;
            hi[end * kQueryBlock] = ri;
            sift(nd, ni, end);
        }
#pragma unroll
        for (int s = 0; s < K; ++s) {
            if (s < nsample) {
                const float dv = __uint_as_float(hd[s * kQueryBlock]);
                idx_row[s] = hi[s * kQueryBlock];
                d2_row[s] = root ? __fsqrt_rn(dv) : dv;
            }
        }
    }
};

// Candidates of a run are requested eight at a time (the offer is a divergent branch, so a load issued next to its use
// would expose its L1/L2 latency on every candidate).
//
// Pending list (round 2, opt-in: see launch_query for the measurement).  A warp executes the 96-instruction list insert whenever ANY of its 32 queries accepts the
// candidate at hand — and with a per-lane acceptance probability of 16 / i for the i-th candidate that is nearly every
// candidate, with 1-4 lanes doing useful work late in the scan (ncu r01q: 14.5 of 32 lanes active over the kernel).
// Here each lane first tests its eight candidates against its current k-th distance and appends the survivors to a
// private pending list (one 64-bit (d2 : idx) word per entry in the lane's own shared-memory column), then drains the
// list: the warp runs max-over-lanes(list length) inserts per eight candidates instead of (almost) eight — ~3 early in
// the scan, 1-2 once the threshold has settled.  Entries that no longer qualify when their turn comes (the threshold moved
// inside the batch) fail the offer's own test.  Same result: the candidate set and the (d2, idx) ranking are unchanged.
constexpr int kPendBatch = 8;

template <class Top>
__device__ __forceinline__ void scan_run(Top &top, const float4 *__restrict__ sorted, int a, int e, float qx,
                                         float qy, float qz, unsigned long long *__restrict__ pend) {
    for (int i = a; i < e; i += kPendBatch) {
        float4 c[kPendBatch];
#pragma unroll
        for (int u = 0; u < kPendBatch; ++u) c[u] = __ldg(sorted + min(i + u, e - 1));
        const float w = top.worst();
        int cnt = 0;
#pragma unroll
        for (int u = 0; u < kPendBatch; ++u) {
            const float d = dist2_ref(qx, qy, qz, c[u].x, c[u].y, c[u].z);
            // '<=': a candidate that ties with the k-th distance can still win on its index (LEX rule)
            if (i + u < e && d <= w) {
                pend[cnt * kQueryBlock] = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_int(c[u].w);
                ++cnt;
            }
        }
        for (int r = 0; r < cnt; ++r) {
            const unsigned long long p = pend[r * kQueryBlock];
            top.offer(__uint_as_float((unsigned)(p >> 32)), (int)(unsigned)p);
        }
    }
}

// the batch-of-four form of round 1 (tuning "knn_pend" = 2): every accepted candidate is inserted where it is found
template <class Top>
__device__ __forceinline__ void scan_run4(Top &top, const float4 *__restrict__ sorted, int a,
                                          int e, float qx, float qy, float qz) {
    int i = a;
    for (; i + 4 <= e; i += 4) {
        const float4 c0 = __ldg(sorted + i), c1 = __ldg(sorted + i + 1), c2 = __ldg(sorted + i + 2),
                     c3 = __ldg(sorted + i + 3);
        const float d0 = dist2_ref(qx, qy, qz, c0.x, c0.y, c0.z), d1 = dist2_ref(qx, qy, qz, c1.x, c1.y, c1.z);
        const float d2 = dist2_ref(qx, qy, qz, c2.x, c2.y, c2.z), d3 = dist2_ref(qx, qy, qz, c3.x, c3.y, c3.z);
        if (fminf(fminf(d0, d1), fminf(d2, d3)) <= top.worst()) {
            top.offer(d0, __float_as_int(c0.w));
            top.offer(d1, __float_as_int(c1.w));
            top.offer(d2, __float_as_int(c2.w));
            top.offer(d3, __float_as_int(c3.w));
        }
    }
    for (; i < e; ++i) {
        const float4 c = __ldg(sorted + i);
        top.offer(dist2_ref(qx, qy, qz, c.x, c.y, c.z), __float_as_int(c.w));
    }
}

template <bool PEND, class Top>
__device__ __forceinline__ void scan(Top &top, const float4 *__restrict__ sorted, int a, int e, float qx, float qy,
                                     float qz, unsigned long long *__restrict__ pend) {
    if (PEND) scan_run(top, sorted, a, e, qx, qy, qz, pend);
    else scan_run4(top, sorted, a, e, qx, qy, qz);
}

template <int K, bool HEAP>
struct TopSel {
    using type = TopK<K, true>;
    static __device__ __forceinline__ void init(type &t) { t.init(); }
    static __device__ __forceinline__ void reset(type &t) { t.init(); }
};
template <int K>
struct TopSel<K, true> {
    using type = HeapK<K>;
    static __device__ __forceinline__ void init(type &t) {
        __shared__ unsigned sd[K * kQueryBlock];
        __shared__ int si[K * kQueryBlock];
        t.init(sd, si);
    }
    static __device__ __forceinline__ void reset(type &t) { t.reset(); }
};

template <int K, bool SELF, bool HEAP, bool PEND>
__global__ void __launch_bounds__(kQueryBlock)
knn_grid_kernel(int m, int b, int nsample, const float *__restrict__ new_xyz,
                const int *__restrict__ new_offset, const GridDesc *__restrict__ desc,
                const int *__restrict__ cells, const float4 *__restrict__ sorted,
                int *__restrict__ idx_out, float *__restrict__ dist2_out, bool root) {
    pdl_wait();
    const int t = blockIdx.x * kQueryBlock + threadIdx.x;
    if (t >= m) return;
    const int sc = find_segment(t, new_offset, b);
    float qx = 0.f, qy = 0.f, qz = 0.f;
    int row = t;  // output row = original query index
    if (SELF) {
        // thread t takes the t-th candidate in cell order (scene ranges are the same before and
        // after the counting sort); points past the last offset belong to no scene and were not sorted
        if (sc < b) {
            const float4 me = __ldg(sorted + t);
            qx = me.x; qy = me.y; qz = me.z;
            row = __float_as_int(me.w);
        }
    } else {
        qx = __ldg(new_xyz + (size_t)t * 3); qy = __ldg(new_xyz + (size_t)t * 3 + 1);
        qz = __ldg(new_xyz + (size_t)t * 3 + 2);
    }
    __shared__ unsigned long long s_pend[PEND ? kPendBatch * kQueryBlock : 1];
    unsigned long long *pend = s_pend + (PEND ? threadIdx.x : 0);
    typename TopSel<K, HEAP>::type top;
    TopSel<K, HEAP>::init(top);
    if (sc < b) {
        const GridDesc g = desc[sc];
        if (g.end > g.start) {
            const int cx = cell_coord(qx, g.ox, g.inv_h, g.nx);
            const int cy = cell_coord(qy, g.oy, g.inv_h, g.ny);
            const int cz = cell_coord(qz, g.oz, g.inv_h, g.nz);
            const int *cs = cells + g.cell_base;
            bool done = false;
            for (int r = 1; r <= kMaxRing && !done; ++r) {
                // rows of the block are visited centre-first (0, -1, +1, -2, +2, ...): the k-th distance drops
                // to its final range within the first rows, so the rows further out offer far fewer candidates
                // that pass the `<= worst` test and trigger the warp-wide insert.  The result does not depend on
                // the visiting order (LEX ranking).
                for (int iz = 0; iz <= 2 * r; ++iz) {
                    const int dz = (iz & 1) ? -((iz + 1) >> 1) : (iz >> 1);
                    const int z = cz + dz;
                    if (z < 0 || z >= g.nz) continue;
                    for (int iy = 0; iy <= 2 * r; ++iy) {
                        const int dy = (iy & 1) ? -((iy + 1) >> 1) : (iy >> 1);
                        const int y = cy + dy;
                        if (y < 0 || y >= g.ny) continue;
                        const int rowbase = (z * g.ny + y) * g.nx;
                        const bool full = (r == 1) || max(abs(dz), abs(dy)) == r;
                        if (full) {
                            const int x0 = max(cx - r, 0), x1 = min(cx + r, g.nx - 1);
                            scan<PEND>(top, sorted, __ldg(cs + rowbase + x0), __ldg(cs + rowbase + x1 + 1), qx, qy, qz, pend);
                        } else {
                            if (cx - r >= 0)
                                scan<PEND>(top, sorted, __ldg(cs + rowbase + cx - r), __ldg(cs + rowbase + cx - r + 1), qx, qy, qz, pend);
                            if (cx + r <= g.nx - 1)
                                scan<PEND>(top, sorted, __ldg(cs + rowbase + cx + r), __ldg(cs + rowbase + cx + r + 1), qx, qy, qz, pend);
                        }
                    }
                }
                // distance to the nearest inner face of the visited block
                float bound = 3.0e38f;
                if (cx - r > 0) bound = fminf(bound, qx - (g.ox + (float)(cx - r) * g.h));
                if (cx + r < g.nx - 1) bound = fminf(bound, (g.ox + (float)(cx + r + 1) * g.h) - qx);
                if (cy - r > 0) bound = fminf(bound, qy - (g.oy + (float)(cy - r) * g.h));
                if (cy + r < g.ny - 1) bound = fminf(bound, (g.oy + (float)(cy + r + 1) * g.h) - qy);
                if (cz - r > 0) bound = fminf(bound, qz - (g.oz + (float)(cz - r) * g.h));
                if (cz + r < g.nz - 1) bound = fminf(bound, (g.oz + (float)(cz + r + 1) * g.h) - qz);
                if (bound > 1.0e38f) {
                    done = true;  // the block covers the whole grid
                } else {
                    const float bs = bound - g.margin;
                    if (bs > 0.f && top.worst() < bs * bs) done = true;
                }
            }
            if (!done) {  // sparse neighbourhood: exhaustive scan of the scene
                TopSel<K, HEAP>::reset(top);
                scan<PEND>(top, sorted, g.start, g.end, qx, qy, qz, pend);
            }
        }
    }
    top.store(idx_out + (size_t)row * nsample, dist2_out + (size_t)row * nsample, nsample, root);
}

// ---- the same search with ONE insert site (tuning "knn_site" = 1, AOPT_KNN_SITE=1; experiment) --------------------------
// ptxas outlines the list insert of knn_grid_kernel — 20 inlined offer() sites: four per batch plus the tail loop, at four scan
// sites — into a subroutine (SASS: 16 CALL.REL.NOINC, one RET, and the STL / LDL pairs that `ptxas -v` reports as a 32-byte
// spill frame whatever the register limit).  Here the traversal produces at most two candidate runs per grid row and ONE loop
// consumes them: batches of four with the tail masked to +inf instead of a tail loop, the accepted batch offered by a
// four-trip loop that is not unrolled, and the exhaustive scan of a sparse neighbourhood folded in as one more "ring" whose
// single run is the whole scene.  Same candidates, same LEX ranking: identical results.
template <int K, bool SELF>
__global__ void __launch_bounds__(kQueryBlock)
knn_grid1_kernel(int m, int b, int nsample, const float *__restrict__ new_xyz,
                 const int *__restrict__ new_offset, const GridDesc *__restrict__ desc,
                 const int *__restrict__ cells, const float4 *__restrict__ sorted,
                 int *__restrict__ idx_out, float *__restrict__ dist2_out, bool root) {
    pdl_wait();
    const int t = blockIdx.x * kQueryBlock + threadIdx.x;
    if (t >= m) return;
    const int sc = find_segment(t, new_offset, b);
    float qx = 0.f, qy = 0.f, qz = 0.f;
    int row = t;
    if (SELF) {
        if (sc < b) {
            const float4 me = __ldg(sorted + t);
            qx = me.x; qy = me.y; qz = me.z;
            row = __float_as_int(me.w);
        }
    } else {
        qx = __ldg(new_xyz + (size_t)t * 3); qy = __ldg(new_xyz + (size_t)t * 3 + 1);
        qz = __ldg(new_xyz + (size_t)t * 3 + 2);
    }
    TopK<K, true> top;
    top.init();
    if (sc < b) {
        const GridDesc g = desc[sc];
        if (g.end > g.start) {
            const int cx = cell_coord(qx, g.ox, g.inv_h, g.nx);
            const int cy = cell_coord(qy, g.oy, g.inv_h, g.ny);
            const int cz = cell_coord(qz, g.oz, g.inv_h, g.nz);
            const int *cs = cells + g.cell_base;
            bool done = false;
            for (int r = 1; r <= kMaxRing + 1 && !done; ++r) {
                const bool all = r > kMaxRing;   // no stop within kMaxRing shells: start over and scan the whole scene
                if (all) top.init();
                const int side = all ? 0 : 2 * r;
                for (int iz = 0; iz <= side; ++iz) {
                    const int dz = (iz & 1) ? -((iz + 1) >> 1) : (iz >> 1);
                    const int z = cz + dz;
                    if (!all && (z < 0 || z >= g.nz)) continue;
                    for (int iy = 0; iy <= side; ++iy) {
                        const int dy = (iy & 1) ? -((iy + 1) >> 1) : (iy >> 1);
                        const int y = cy + dy;
                        if (!all && (y < 0 || y >= g.ny)) continue;
                        int a0 = 0, e0 = 0, a1 = 0, e1 = 0;
                        if (all) {
                            a0 = g.start; e0 = g.end;
                        } else {
                            const int rowbase = (z * g.ny + y) * g.nx;
                            const bool full = (r == 1) || max(abs(dz), abs(dy)) == r;
                            if (full) {
                                const int x0 = max(cx - r, 0), x1 = min(cx + r, g.nx - 1);
                                a0 = __ldg(cs + rowbase + x0); e0 = __ldg(cs + rowbase + x1 + 1);
                            } else {
                                if (cx - r >= 0) { a0 = __ldg(cs + rowbase + cx - r); e0 = __ldg(cs + rowbase + cx - r + 1); }
                                if (cx + r <= g.nx - 1) { a1 = __ldg(cs + rowbase + cx + r); e1 = __ldg(cs + rowbase + cx + r + 1); }
                            }
                        }
#pragma unroll 1
                        for (int rep = 0; rep < 2; ++rep) {
                            const int a = rep ? a1 : a0, e = rep ? e1 : e0;
                            for (int i = a; i < e; i += 4) {
                                const float4 c0 = __ldg(sorted + i), c1 = __ldg(sorted + min(i + 1, e - 1)),
                                             c2 = __ldg(sorted + min(i + 2, e - 1)), c3 = __ldg(sorted + min(i + 3, e - 1));
                                const float d0 = dist2_ref(qx, qy, qz, c0.x, c0.y, c0.z);
                                const float d1 = i + 1 < e ? dist2_ref(qx, qy, qz, c1.x, c1.y, c1.z) : 3.0e38f;
                                const float d2 = i + 2 < e ? dist2_ref(qx, qy, qz, c2.x, c2.y, c2.z) : 3.0e38f;
                                const float d3 = i + 3 < e ? dist2_ref(qx, qy, qz, c3.x, c3.y, c3.z) : 3.0e38f;
                                if (fminf(fminf(d0, d1), fminf(d2, d3)) <= top.worst()) {
#pragma unroll 1
                                    for (int u = 0; u < 4; ++u) {
                                        const float du = u == 0 ? d0 : u == 1 ? d1 : u == 2 ? d2 : d3;
                                        const float wu = u == 0 ? c0.w : u == 1 ? c1.w : u == 2 ? c2.w : c3.w;
                                        top.offer(du, __float_as_int(wu));   // masked slots carry 3e38: never accepted
                                    }
                                }
                            }
                        }
                    }
                }
                if (all) break;
                float bound = 3.0e38f;
                if (cx - r > 0) bound = fminf(bound, qx - (g.ox + (float)(cx - r) * g.h));
                if (cx + r < g.nx - 1) bound = fminf(bound, (g.ox + (float)(cx + r + 1) * g.h) - qx);
                if (cy - r > 0) bound = fminf(bound, qy - (g.oy + (float)(cy - r) * g.h));
                if (cy + r < g.ny - 1) bound = fminf(bound, (g.oy + (float)(cy + r + 1) * g.h) - qy);
                if (cz - r > 0) bound = fminf(bound, qz - (g.oz + (float)(cz - r) * g.h));
                if (cz + r < g.nz - 1) bound = fminf(bound, (g.oz + (float)(cz + r + 1) * g.h) - qz);
                if (bound > 1.0e38f) {
                    done = true;
                } else {
                    const float bs = bound - g.margin;
                    if (bs > 0.f && top.worst() < bs * bs) done = true;
                }
            }
        }
    }
    top.store(idx_out + (size_t)row * nsample, dist2_out + (size_t)row * nsample, nsample, root);
}

// ---- host side -------------------------------------------------------------------------------------
static size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }

struct GridWs {
    GridDesc *desc;
    float *samples;
    unsigned *bbox;  // (b,3) encoded minima then (b,3) encoded maxima
    int *cells;
    int *point_cell, *point_slot;
    float4 *sorted;
    int *partial;
    size_t total_cells;
    size_t bytes;
};

static GridWs carve(void *ws, int n, int b) {
    GridWs w;
    char *p = static_cast<char *>(ws);
    size_t off = 0;
    w.total_cells = (size_t)cells_per_point() * n + (size_t)kCellsPerScene * b;
    w.desc = reinterpret_cast<GridDesc *>(p + off); off += a256(sizeof(GridDesc) * (size_t)(b > 0 ? b : 1));
    w.samples = reinterpret_cast<float *>(p + off); off += a256(4 * (size_t)kSamples * (b > 0 ? b : 1));
    w.bbox = reinterpret_cast<unsigned *>(p + off); off += a256(4 * 6 * (size_t)(b > 0 ? b : 1));
    w.cells = reinterpret_cast<int *>(p + off); off += a256(4 * (w.total_cells + 1));
    w.point_cell = reinterpret_cast<int *>(p + off); off += a256(4 * (size_t)n);
    w.point_slot = reinterpret_cast<int *>(p + off); off += a256(4 * (size_t)n);
    w.sorted = reinterpret_cast<float4 *>(p + off); off += a256(16 * (size_t)n);
    w.partial = reinterpret_cast<int *>(p + off); off += a256(4 * scan_partial_ints((long long)w.total_cells));
    w.bytes = off;
    return w;
}

size_t knn_grid_workspace_bytes(int n, int m, int b) {
    (void)m;
    return carve(nullptr, n, b).bytes;
}

template <int K, bool HEAP, bool PEND>
static void launch_query_t(bool self, int m, int b, int nsample, const float *new_xyz, const int *new_offset,
                           const GridWs &w, int *idx, float *dist2, bool root, cudaStream_t st) {
    const int grid = div_up(m, kQueryBlock);
    const bool pdl = tuning(kTunePdl) != 2;
    if (self)
        launch_chain(pdl, knn_grid_kernel<K, true, HEAP, PEND>, grid, kQueryBlock, 0, st, m, b, nsample, new_xyz, new_offset,
                     (const GridDesc *)w.desc, (const int *)w.cells, (const float4 *)w.sorted, idx, dist2, root);
    else
        launch_chain(pdl, knn_grid_kernel<K, false, HEAP, PEND>, grid, kQueryBlock, 0, st, m, b, nsample, new_xyz, new_offset,
                     (const GridDesc *)w.desc, (const int *)w.cells, (const float4 *)w.sorted, idx, dist2, root);
}

// Container choice (tuning "knn_topk", AOPT_KNN_TOPK=heap|list).  Default: the register list.  Measured on a B200
// (profiles/r02m_knn_topk_ab.txt, whole search, identical bits): the heap wins only at k = 32 on the dense indoor
// scans (S3DIS level 0: 515 vs 634 us, ScanNet 3 x 150k: 776 vs 967 us); at k = 16 it is level (317 vs 313 us at
// level 0) or slower (173 vs 146 us at level 1, 113 vs 84 us at level 3), at k = 8 slower everywhere: a sift is
// <= 4 dependent shared-memory round trips with the lanes at different depths, which costs the warp as many issue
// slots as the 96 independent compares / selects of the list insert.
// Pending list (tuning "knn_pend", AOPT_KNN_PEND=1|0), K >= 8.  Default: off.  Measured (profiles/r02s_knn_variants_ab.txt,
// identical bits): level 0, k = 16: 335 us with the list, 319 us inserting in place; level 1: 180 vs 147 us; ScanNet
// 3 x 150k: 430 vs 385 us; it gains only at k = 32 (KITTI 811 vs 891 us).  The premise — a warp inserting for 1-4 of its
// lanes through most of the scan — does not hold at k = 16: a query sees only ~60 candidates in its 27 cells (4 cells per
// point: most cells of an indoor scan are empty), so the acceptance probability 16 / i stays above 0.25 to the end and
// the longest pending list of a warp is as long as the batch.
template <int K>
static void launch_query(bool self, int m, int b, int nsample, const float *new_xyz, const int *new_offset,
                         const GridWs &w, int *idx, float *dist2, bool root, cudaStream_t st) {
    if (tuning(kTuneKnnSite) == 1 && tuning(kTuneKnnTopk) != 1 && tuning(kTuneKnnPend) != 1) {
        const int grid = div_up(m, kQueryBlock);
        const bool pdl = tuning(kTunePdl) != 2;
        if (self)
            launch_chain(pdl, knn_grid1_kernel<K, true>, grid, kQueryBlock, 0, st, m, b, nsample, new_xyz, new_offset,
                         (const GridDesc *)w.desc, (const int *)w.cells, (const float4 *)w.sorted, idx, dist2, root);
        else
            launch_chain(pdl, knn_grid1_kernel<K, false>, grid, kQueryBlock, 0, st, m, b, nsample, new_xyz, new_offset,
                         (const GridDesc *)w.desc, (const int *)w.cells, (const float4 *)w.sorted, idx, dist2, root);
        return;
    }
    if constexpr (K >= 8) {
        const bool pend = tuning(kTuneKnnPend) == 1;
        if (tuning(kTuneKnnTopk) == 1) {
            if (pend) launch_query_t<K, true, true>(self, m, b, nsample, new_xyz, new_offset, w, idx, dist2, root, st);
            else launch_query_t<K, true, false>(self, m, b, nsample, new_xyz, new_offset, w, idx, dist2, root, st);
            return;
        }
        if (pend) {
            launch_query_t<K, false, true>(self, m, b, nsample, new_xyz, new_offset, w, idx, dist2, root, st);
            return;
        }
    }
    launch_query_t<K, false, false>(self, m, b, nsample, new_xyz, new_offset, w, idx, dist2, root, st);
}

template <int K>
static void launch_sample(int b, int nsample, const float *xyz, const int *offset, const GridWs &w, cudaStream_t st) {
    knn_sample_kernel<K><<<b * kSamples, kSampleBlock, 0, st>>>(b, nsample, xyz, offset, w.samples, w.cells,
                                                                (long long)w.total_cells + 1, w.bbox, w.bbox + 3 * (size_t)b);
}

int knn_grid_launch(int m, int nsample, int n, int b, const float *xyz, const float *new_xyz,
                    const int *offset, const int *new_offset, int *idx, float *dist2, bool root, void *ws,
                    size_t ws_bytes, cudaStream_t st) {
    if (b <= 0 || n <= 0) {  // no candidates: everything is padding
        return AOPT_ERR_INVALID_ARGUMENT;
    }
    GridWs w = carve(ws, n, b);
    if (ws_bytes < w.bytes) return AOPT_ERR_WORKSPACE;
    float scale = 1.25f;
    if (const char *e = getenv("AOPT_KNN_CELL_SCALE")) {
        float v = (float)atof(e);
        if (v > 0.01f && v < 100.f) scale = v;
    }
    const bool self = (new_xyz == xyz) && (new_offset == offset) && (m == n);

    const bool sampled = tuning(kTuneKnnSample) != 1;
    if (!sampled) {
        const long long nc = (long long)w.total_cells + 1;
        knn_clear_kernel<<<stride_grid(nc, 256, 8), 256, 0, st>>>(b, w.cells, nc, w.bbox, w.bbox + 3 * (size_t)b);
    }
    else if (nsample <= 1) launch_sample<1>(b, nsample, xyz, offset, w, st);
    else if (nsample <= 3) launch_sample<3>(b, nsample, xyz, offset, w, st);
    else if (nsample <= 4) launch_sample<4>(b, nsample, xyz, offset, w, st);
    else if (nsample <= 8) launch_sample<8>(b, nsample, xyz, offset, w, st);
    else if (nsample <= 16) launch_sample<16>(b, nsample, xyz, offset, w, st);
    else launch_sample<32>(b, nsample, xyz, offset, w, st);
    // sample -> bbox -> setup -> count -> scan -> fill -> query: one chain of programmatic dependent launches
    const bool pdl = tuning(kTunePdl) != 2;
    launch_scene_bbox(n, b, xyz, offset, w.bbox, w.bbox + 3 * (size_t)b, st, /*init=*/false, pdl);
    launch_chain(pdl, grid_setup_kernel, div_up(b, 4), 128, 0, st, b, n, offset, (const unsigned *)w.bbox,
                 (const unsigned *)(w.bbox + 3 * (size_t)b), (const float *)(sampled ? w.samples : nullptr), scale,
                 cells_per_point(), w.desc);
    launch_chain(pdl, grid_count_kernel, div_up(n, 256), 256, 0, st, n, b, xyz, offset, (const GridDesc *)w.desc, w.cells,
                 w.point_cell, w.point_slot, w.partial, (int)scan_partial_ints((long long)w.total_cells));
    launch_exclusive_scan_chained(w.cells, w.cells, (int)w.total_cells, w.partial, st, pdl);
    launch_chain(pdl, grid_fill_kernel, div_up(n, 256), 256, 0, st, n, xyz, (const int *)w.cells, (const int *)w.point_cell,
                 (const int *)w.point_slot, w.sorted);
    if (nsample <= 1) launch_query<1>(self, m, b, nsample, new_xyz, new_offset, w, idx, dist2, root, st);
    else if (nsample <= 3) launch_query<3>(self, m, b, nsample, new_xyz, new_offset, w, idx, dist2, root, st);
    else if (nsample <= 4) launch_query<4>(self, m, b, nsample, new_xyz, new_offset, w, idx, dist2, root, st);
    else if (nsample <= 8) launch_query<8>(self, m, b, nsample, new_xyz, new_offset, w, idx, dist2, root, st);
    else if (nsample <= 16) launch_query<16>(self, m, b, nsample, new_xyz, new_offset, w, idx, dist2, root, st);
    else launch_query<32>(self, m, b, nsample, new_xyz, new_offset, w, idx, dist2, root, st);
    return check_launch(7);  // sample, bbox, setup, count, scan, fill, query
}

}  // namespace aopt
