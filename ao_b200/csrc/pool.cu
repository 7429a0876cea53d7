// pool.cu — GridPool: voxel keys and the per-voxel segment reduction (mean coord, max feature).
//
// Replaces, from /root/reference/pointcept/models/point_transformer_v2/point_transformer_v2m2_base.py:
//   :249-253  segment_csr(coord, ptr, "min")            → aopt_segment_min3
//   :257-259  voxel_grid(coord - start[batch], size, batch, start=0)   → aopt_voxel_keys
//   :265-266  segment_csr(coord[sorted], idx_ptr, "mean"), segment_csr(feat[sorted], idx_ptr, "max")
//                                                        → aopt_pool_forward (no permuted copies)
//   autograd of :266 (scatter of the gradient to the arg-max rows)     → aopt_pool_backward
// torch_scatter / torch_cluster are not vendored by the reference; their semantics are restated
// in oracle/torch_ref.py (segment_csr, voxel_grid_keys) and these kernels follow that restatement.
//
// Bytes: forward reads 4NC + 12N + 4N (order) + 4(N'+1), writes 8N'C + 12N';
// backward reads 8N'C + 4N, writes 4NC.
#include <float.h>

#include "common.cuh"
#include "knn_common.cuh"  // find_segment
#include "scan.cuh"
#include "bbox.cuh"
#include "radix.cuh"

namespace aopt {

constexpr int kPoolBlock = 256;

// start[sc, a] = minimum coordinate of scene sc (0 for an empty scene, like segment_csr's fill value):
// decode of the integer-encoded minima accumulated in place by scene_bbox_kernel (bbox.cu).
__global__ void __launch_bounds__(128)
decode_min_kernel(int count, float *__restrict__ start) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= count) return;
    const unsigned e = reinterpret_cast<unsigned *>(start)[i];
    start[i] = e == kBboxEmptyLo ? 0.f : bbox_decode(e);
}

// fixed layout of the stand-alone key (aopt_voxel_keys): 3 x 18 cell bits + 9 scene bits = 63: the key stays
// non-negative, so a signed 64-bit sort keeps the scene as the most significant digit (a 10th scene bit would be the
// sign bit).  aopt_voxel_grid below packs its own, compact layout and has no such limit.
constexpr int kCellBits = 18, kSceneBits = 9;

__global__ void __launch_bounds__(kPoolBlock)
voxel_keys_kernel(int n, int b, const float *__restrict__ coord, const int *__restrict__ offset,
                  const float *__restrict__ start, float grid_size, int64_t *__restrict__ keys,
                  int *__restrict__ status_flag) {
    const int i = blockIdx.x * kPoolBlock + threadIdx.x;
    if (i >= n) return;
    int sc = find_segment(i, offset, b);
    if (sc >= b) sc = b - 1;
    int64_t cell[3];
    bool bad = sc >= (1 << kSceneBits);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        // fp32 subtract then IEEE fp32 divide then truncate — the arithmetic of
        // (coord - start[batch]) / size → .long() in torch_cluster.grid_cluster
        float rel = __fsub_rn(__ldg(coord + (size_t)i * 3 + a), __ldg(start + sc * 3 + a));
        float q = __fdiv_rn(rel, grid_size);
        int64_t cq = (int64_t)q;
        if (cq < 0 || cq >= (1LL << kCellBits)) { bad = true; cq = cq < 0 ? 0 : (1LL << kCellBits) - 1; }
        cell[a] = cq;
    }
    keys[i] = ((int64_t)sc << (3 * kCellBits)) | (cell[2] << (2 * kCellBits)) | (cell[1] << kCellBits) | cell[0];
    if (bad && status_flag) *status_flag = 1;
}

// One thread per (voxel, 4-channel chunk): running max + the original id of the first maximal point.
// The voxel walk is a dependent chain (idx_ptr → order → feat row), so the 128-bit path takes the points
// of a voxel eight at a time: eight `order` loads, then eight feature pieces requested with cp.async into
// the thread's own shared-memory slots, the next eight `order` values prefetched meanwhile.
constexpr int kPoolBatch = 8;

template <int VEC>
__global__ void __launch_bounds__(kPoolBlock)
pool_forward_kernel(long long n_vox, int chunks, int c, const float *__restrict__ feat,
                    const float *__restrict__ coord, const int *__restrict__ order,
                    const int *__restrict__ idx_ptr, float *__restrict__ out_feat,
                    int *__restrict__ argmax, float *__restrict__ out_coord) {
    __shared__ float4 stage[VEC == 4 ? kPoolBatch * kPoolBlock : 1];
    float4 *sf = stage + threadIdx.x;
    const ColWalk cw = col_walk(chunks, kPoolBlock);  // the thread keeps its chunk: no division per item
    const int col = cw.col;
    for (long long v = cw.row; v < n_vox; v += cw.row_step) {
        const int e0 = __ldg(idx_ptr + v), e1 = __ldg(idx_ptr + v + 1);
        float best[VEC];
        int arg[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) { best[i] = -FLT_MAX; arg[i] = -1; }
        float sx = 0.f, sy = 0.f, sz = 0.f;
        const bool do_coord = col == 0 && out_coord;
        if constexpr (VEC == 4) {
            int pn[kPoolBatch];
#pragma unroll
            for (int u = 0; u < kPoolBatch; ++u) pn[u] = (e0 + u < e1) ? __ldg(order + e0 + u) : -1;
            for (int e = e0; e < e1; e += kPoolBatch) {
                int pt[kPoolBatch];
#pragma unroll
                for (int u = 0; u < kPoolBatch; ++u) {
                    // slots past the end of the voxel (most voxels hold fewer than 8 points) re-request the
                    // batch's first point — a line this thread already asked for.  They used to request point 0:
                    // every thread of the grid hammering ONE L2 line (the csr_walk.cuh hot-spot rule).
                    pt[u] = pn[u] >= 0 ? pn[u] : pn[0];
                    cp_async16_stream(sf + u * kPoolBlock, feat + (size_t)pt[u] * c + col * 4);
                }
                cp_async_commit();
#pragma unroll
                for (int u = 0; u < kPoolBatch; ++u)
                    pn[u] = (e + kPoolBatch + u < e1) ? __ldg(order + e + kPoolBatch + u) : -1;
                if (do_coord) {  // sequential sum in `order`, like segment_csr
#pragma unroll
                    for (int u = 0; u < kPoolBatch; ++u) {
                        if (e + u < e1) {
                            sx += __ldg(coord + (size_t)pt[u] * 3 + 0);
                            sy += __ldg(coord + (size_t)pt[u] * 3 + 1);
                            sz += __ldg(coord + (size_t)pt[u] * 3 + 2);
                        }
                    }
                }
                cp_async_wait_all();
#pragma unroll
                for (int u = 0; u < kPoolBatch; ++u) {
                    if (e + u < e1) {
                        const float4 q = sf[u * kPoolBlock];
                        const float x[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (x[i] > best[i]) { best[i] = x[i]; arg[i] = pt[u]; }
                    }
                }
            }
        } else {
            for (int e = e0; e < e1; ++e) {
                const int pt = __ldg(order + e);
                const float x = __ldg(feat + (size_t)pt * c + col);
                if (x > best[0]) { best[0] = x; arg[0] = pt; }
                if (do_coord) {
                    sx += __ldg(coord + (size_t)pt * 3 + 0);
                    sy += __ldg(coord + (size_t)pt * 3 + 1);
                    sz += __ldg(coord + (size_t)pt * 3 + 2);
                }
            }
        }
        const size_t o = (size_t)v * c + col * VEC;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            if (arg[i] < 0) best[i] = 0.f;  // nothing compared greater (empty / all NaN): segment_csr writes 0
        }
        if constexpr (VEC == 4) {
            *reinterpret_cast<float4 *>(out_feat + o) = make_float4(best[0], best[1], best[2], best[3]);
            *reinterpret_cast<int4 *>(argmax + o) = make_int4(arg[0], arg[1], arg[2], arg[3]);
        } else {
            out_feat[o] = best[0];
            argmax[o] = arg[0];
        }
        if (do_coord) {
            const float cnt = (float)max(e1 - e0, 1);
            out_coord[v * 3 + 0] = sx / cnt;
            out_coord[v * 3 + 1] = sy / cnt;
            out_coord[v * 3 + 2] = sz / cnt;
        }
    }
}

// Gather form of the max-pool backward: every (point, channel) looks up its own voxel, so the
// whole (N,C) gradient is written once, coalesced, with no memset and no atomics.
template <int VEC>
__global__ void __launch_bounds__(kPoolBlock)
pool_backward_kernel(long long n, int chunks, int c, const float *__restrict__ grad_out,
                     const int *__restrict__ argmax, const int *__restrict__ cluster,
                     float *__restrict__ grad_feat) {
    const ColWalk cw = col_walk(chunks, kPoolBlock);  // the thread keeps its chunk: no division per item
    const int col = cw.col;
    for (long long pt = cw.row; pt < n; pt += cw.row_step) {
        const int v = __ldg(cluster + pt);
        const size_t src = (size_t)v * c + col * VEC, dst = (size_t)pt * c + col * VEC;
        if constexpr (VEC == 4) {
            const int4 a = __ldg(reinterpret_cast<const int4 *>(argmax + src));
            const float4 gq = ldg_gather4(grad_out + src);
            float4 r;
            r.x = a.x == (int)pt ? gq.x : 0.f; r.y = a.y == (int)pt ? gq.y : 0.f;
            r.z = a.z == (int)pt ? gq.z : 0.f; r.w = a.w == (int)pt ? gq.w : 0.f;
            stg_stream4(grad_feat + dst, r);
        } else {
            grad_feat[dst] = __ldg(argmax + src) == (int)pt ? __ldg(grad_out + src) : 0.f;
        }
    }
}

// ---- voxel partition from sorted keys (the tail of …v2m2_base.py:260-268 without torch.unique) -------
// flag[i] = 1 where a new voxel starts in the sorted key sequence.
__global__ void __launch_bounds__(kPoolBlock)
voxel_mark_kernel(int n, const int64_t *__restrict__ sorted_keys, int *__restrict__ flag) {
    const int i = blockIdx.x * kPoolBlock + threadIdx.x;
    if (i >= n) return;
    flag[i] = (i == 0 || __ldg(sorted_keys + i) != __ldg(sorted_keys + i - 1)) ? 1 : 0;
}

// scan[i] = exclusive prefix sum of flag → voxel id of sorted position i is scan[i] + flag[i] - 1.
// Writes order32, cluster (both widths), idx_ptr, the per-scene voxel offsets and meta = {n_vox}.
__global__ void __launch_bounds__(kPoolBlock)
voxel_finalize_kernel(int n, int b, const int64_t *__restrict__ order64, const int *__restrict__ flag,
                      const int *__restrict__ scan, const int *__restrict__ offset,
                      int *__restrict__ order32, int *__restrict__ cluster32, int64_t *__restrict__ cluster64,
                      int *__restrict__ idx_ptr, int64_t *__restrict__ new_offset, int *__restrict__ meta) {
    const int i = blockIdx.x * kPoolBlock + threadIdx.x;
    const int n_vox = __ldg(scan + n);  // total number of flags
    if (i < n) {
        const int f = __ldg(flag + i);
        const int vid = __ldg(scan + i) + f - 1;
        const int pt = (int)__ldg(order64 + i);
        order32[i] = pt;
        cluster32[pt] = vid;
        cluster64[pt] = vid;
        if (f) idx_ptr[vid] = i;
    }
    if (i == 0) {
        idx_ptr[n_vox] = n;
        meta[0] = n_vox;
    }
    // scene ranges are identical before and after the sort (the scene id is the most significant key
    // digit), so the voxels of scenes 0..s are those starting before sorted position offset[s]
    if (i < b) {
        int e = min(max(__ldg(offset + i), 0), n);
        new_offset[i] = e > 0 ? (int64_t)(__ldg(scan + e - 1) + __ldg(flag + e - 1)) : 0;
    }
}

// ---- voxel partition, all on the device: bbox -> compact keys -> own radix sort -> partition ------------------------
// (…v2m2_base.py:246-268: offset2batch, segment_csr(min), voxel_grid, torch.unique, torch.sort in one entry point.)
// The fixed 3 x 18 + 10 bit key above needs a 64-bit sort (eight library passes).  The cell extents of the batch are
// known once the per-scene bounding boxes are: the key is packed into the FEWEST bits that hold (scene, z, y, x) —
// 22 bits for four S3DIS rooms at 0.1 m, 25 for two KITTI scans at 0.15 m — and sorted by radix.cuh in
// ceil(bits / 11) passes.  The widths live in a device word (meta[2] = passes needed); the host always enqueues
// `max_passes` passes and the unneeded ones return at once, so there is no extra host synchronisation.
// Order of the keys = (scene, z, y, x) ascending = the order torch.unique(sorted=True) gives grid_cluster's keys.
// meta: [0] number of voxels, [1] flags (1: a cell index is negative or the key needs more than 64 bits,
//       2: more passes needed than were enqueued), [2] passes needed, [3] key bits, [4..6] shifts of y, z, scene.
constexpr int kMetaInts = 8;

__global__ void __launch_bounds__(128)
voxel_layout_kernel(int b, const unsigned *__restrict__ lo, const unsigned *__restrict__ hi,
                    const float *__restrict__ start_in, float grid_size, int max_passes,
                    float *__restrict__ start, int *__restrict__ meta) {
    pdl_wait();
    pdl_trigger();
    __shared__ long long cell_max[3];
    if (threadIdx.x < 3) cell_max[threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * b; i += 128) {
        const unsigned elo = lo[i], ehi = hi[i];
        // an empty scene has no minimum: 0, like segment_csr's fill value
        const float st = start_in ? start_in[i] : (elo == kBboxEmptyLo ? 0.f : bbox_decode(elo));
        start[i] = st;
        if (ehi != kBboxEmptyHi || elo != kBboxEmptyLo) {
            // fsub / fdiv are monotone in the coordinate, so the scene's largest coordinate has its largest cell
            const float q = __fdiv_rn(__fsub_rn(bbox_decode(ehi), st), grid_size);
            long long cq = (long long)q;
            if (!(q < 9.0e18f)) cq = 0x7fffffffffffffffLL;
            if (cq > 0) atomicMax(reinterpret_cast<unsigned long long *>(&cell_max[i % 3]), (unsigned long long)cq);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        auto bits_of = [](unsigned long long v) { int n = 0; while (v) { ++n; v >>= 1; } return n; };
        const int bx = bits_of((unsigned long long)cell_max[0]), by = bits_of((unsigned long long)cell_max[1]),
                  bz = bits_of((unsigned long long)cell_max[2]), bs = bits_of((unsigned long long)(b - 1));
        const int total = bx + by + bz + bs;
        int npass = (total + kRadixBits - 1) / kRadixBits;
        if (npass < 1) npass = 1;
        int flags = 0;
        if (total > 64) flags |= 1;
        if (npass > max_passes) flags |= 2;
        meta[0] = 0; meta[1] = flags; meta[2] = npass; meta[3] = total;
        meta[4] = bx; meta[5] = bx + by; meta[6] = bx + by + bz; meta[7] = 0;
    }
}

__global__ void __launch_bounds__(kPoolBlock)
voxel_ckeys_kernel(int n, int b, const float *__restrict__ coord, const int *__restrict__ offset,
                   const float *__restrict__ start, float grid_size, unsigned long long *__restrict__ keys,
                   int *__restrict__ meta) {
    pdl_wait();
    pdl_trigger();
    const int i = blockIdx.x * kPoolBlock + threadIdx.x;
    if (i >= n) return;
    int sc = find_segment(i, offset, b);
    if (sc >= b) sc = b - 1;
    const int sy = meta[4], sz = meta[5], ss = meta[6];
    unsigned long long cell[3];
    bool bad = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        // fp32 subtract then IEEE fp32 divide then truncate — the arithmetic of
        // (coord - start[batch]) / size -> .long() in torch_cluster.grid_cluster
        const float rel = __fsub_rn(__ldg(coord + (size_t)i * 3 + a), __ldg(start + sc * 3 + a));
        const long long cq = (long long)__fdiv_rn(rel, grid_size);
        if (cq < 0) bad = true;   // only with a caller-provided start above the scene minimum
        cell[a] = cq < 0 ? 0ull : (unsigned long long)cq;
    }
    // a cell beyond the widths (possible only next to a `bad` one) must not spill into the next field
    const unsigned long long mx = sy >= 64 ? ~0ull : ((1ull << sy) - 1), my = (sz - sy) >= 64 ? ~0ull : ((1ull << (sz - sy)) - 1),
                             mz = (ss - sz) >= 64 ? ~0ull : ((1ull << (ss - sz)) - 1);
    if (cell[0] > mx || cell[1] > my || cell[2] > mz) bad = true;
    keys[i] = (ss < 64 ? ((unsigned long long)sc << ss) : 0ull) | (sz < 64 ? ((cell[2] & mz) << sz) : 0ull) |
              (sy < 64 ? ((cell[1] & my) << sy) : 0ull) | (cell[0] & mx);
    if (bad) atomicOr(meta + 1, 1);
}

// result of the sort: buffer P after an odd number of passes, Q after an even number
__global__ void __launch_bounds__(kPoolBlock)
voxel_mark2_kernel(int n, const unsigned long long *__restrict__ kp, const unsigned long long *__restrict__ kq,
                   const int *__restrict__ meta, int *__restrict__ flag, int *__restrict__ scan_state, int scan_state_ints) {
    pdl_wait();
    pdl_trigger();
    for (int t = blockIdx.x * kPoolBlock + threadIdx.x; t < scan_state_ints; t += gridDim.x * kPoolBlock) scan_state[t] = 0;
    const int i = blockIdx.x * kPoolBlock + threadIdx.x;
    if (i >= n) return;
    const unsigned long long *sorted_keys = (meta[2] & 1) ? kp : kq;
    flag[i] = (i == 0 || sorted_keys[i] != sorted_keys[i - 1]) ? 1 : 0;
}

__global__ void __launch_bounds__(kPoolBlock)
voxel_finalize2_kernel(int n, int b, const int *__restrict__ vp, const int *__restrict__ vq,
                       const int *__restrict__ flag, const int *__restrict__ scan, const int *__restrict__ offset,
                       int *__restrict__ order32, int *__restrict__ cluster32, int64_t *__restrict__ cluster64,
                       int *__restrict__ idx_ptr, int64_t *__restrict__ new_offset, int *__restrict__ meta) {
    pdl_wait();
    const int i = blockIdx.x * kPoolBlock + threadIdx.x;
    const int n_vox = __ldg(scan + n);
    const int *order = (meta[2] & 1) ? vp : vq;
    if (i < n) {
        const int f = __ldg(flag + i);
        const int vid = __ldg(scan + i) + f - 1;
        const int pt = order[i];
        order32[i] = pt;
        cluster32[pt] = vid;
        cluster64[pt] = vid;
        if (f) idx_ptr[vid] = i;
    }
    if (i == 0) {
        idx_ptr[n_vox] = n;
        meta[0] = n_vox;
    }
    // scene ranges are identical before and after the sort (the scene id is the most significant key field)
    if (i < b) {
        int e = min(max(__ldg(offset + i), 0), n);
        new_offset[i] = e > 0 ? (int64_t)(__ldg(scan + e - 1) + __ldg(flag + e - 1)) : 0;
    }
}

__global__ void __launch_bounds__(256)
offset2batch_kernel(int n, int b, const int *__restrict__ offset, int64_t *__restrict__ batch) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) batch[i] = find_segment(i, offset, b);
}

}  // namespace aopt

using namespace aopt;

extern "C" int aopt_offset2batch(int n, int b, const int *offset, int64_t *batch, aopt_stream_t stream) {
    if (n < 0 || b < 0) return AOPT_ERR_INVALID_ARGUMENT;
    if (n == 0) return AOPT_OK;
    if (!offset || !batch) return AOPT_ERR_INVALID_ARGUMENT;
    offset2batch_kernel<<<div_up(n, 256), 256, 0, as_stream(stream)>>>(n, b, offset, batch);
    return check_launch();
}

extern "C" int aopt_segment_min3(int n, int b, const float *coord, const int *offset, float *start,
                                 aopt_stream_t stream) {
    if (n < 0 || b < 0) return AOPT_ERR_INVALID_ARGUMENT;
    if (b == 0) return AOPT_OK;
    if (!coord || !offset || !start) return AOPT_ERR_INVALID_ARGUMENT;
    launch_scene_bbox(n, b, coord, offset, reinterpret_cast<unsigned *>(start), nullptr, as_stream(stream));
    decode_min_kernel<<<div_up(3 * b, 128), 128, 0, as_stream(stream)>>>(3 * b, start);
    return check_launch(2);
}

extern "C" int aopt_voxel_keys(int n, int b, const float *coord, const int *offset, const float *start,
                               float grid_size, int64_t *keys, int *status_flag, aopt_stream_t stream) {
    if (n < 0 || b < 1 || !(grid_size > 0.f)) return AOPT_ERR_INVALID_ARGUMENT;
    if (n == 0) return AOPT_OK;
    if (!coord || !offset || !start || !keys) return AOPT_ERR_INVALID_ARGUMENT;
    voxel_keys_kernel<<<div_up(n, kPoolBlock), kPoolBlock, 0, as_stream(stream)>>>(n, b, coord, offset, start,
                                                                               grid_size, keys, status_flag);
    return check_launch();
}

extern "C" size_t aopt_voxel_partition_workspace_bytes(int n) {
    if (n < 0) return 0;
    auto a256 = [](size_t x) { return (x + 255) & ~(size_t)255; };
    return a256(4 * ((size_t)n + 1)) * 2 + a256(4 * scan_partial_ints(n));
}

extern "C" int aopt_voxel_partition(int n, int b, const int64_t *sorted_keys, const int64_t *order64,
                                    const int *offset, int *order32, int *cluster32, int64_t *cluster64,
                                    int *idx_ptr, int64_t *new_offset, int *meta, void *workspace,
                                    size_t workspace_bytes, aopt_stream_t stream) {
    if (n < 1 || b < 1) return AOPT_ERR_INVALID_ARGUMENT;
    if (!sorted_keys || !order64 || !offset || !order32 || !cluster32 || !cluster64 || !idx_ptr || !new_offset || !meta)
        return AOPT_ERR_INVALID_ARGUMENT;
    if (!workspace || workspace_bytes < aopt_voxel_partition_workspace_bytes(n)) return AOPT_ERR_WORKSPACE;
    auto a256 = [](size_t x) { return (x + 255) & ~(size_t)255; };
    char *ws = static_cast<char *>(workspace);
    int *flag = reinterpret_cast<int *>(ws); ws += a256(4 * ((size_t)n + 1));
    int *scan = reinterpret_cast<int *>(ws); ws += a256(4 * ((size_t)n + 1));
    int *partial = reinterpret_cast<int *>(ws);
    cudaStream_t st = as_stream(stream);
    const int grid = div_up(n > b ? n : b, kPoolBlock);
    voxel_mark_kernel<<<div_up(n, kPoolBlock), kPoolBlock, 0, st>>>(n, sorted_keys, flag);
    launch_exclusive_scan(flag, scan, n, partial, st);
    voxel_finalize_kernel<<<grid, kPoolBlock, 0, st>>>(n, b, order64, flag, scan, offset, order32, cluster32, cluster64,
                                                       idx_ptr, new_offset, meta);
    return check_launch(3);  // mark, scan, finalize
}

static size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t aopt_voxel_grid_workspace_bytes(int n, int b) {
    if (n < 0 || b < 0) return 0;
    return 3 * a256(12 * (size_t)b) + 3 * a256(8 * (size_t)n) + 2 * a256(4 * (size_t)n) + 2 * a256(4 * ((size_t)n + 1)) +
           a256(4 * radix_scratch_ints(n) + 16) + a256(4 * scan_partial_ints(n));
}

extern "C" int aopt_voxel_grid(int n, int b, const float *coord, const int *offset, const float *start,
                               float grid_size, int max_passes, int *order32, int *cluster32, int64_t *cluster64,
                               int *idx_ptr, int64_t *new_offset, int *meta, void *workspace,
                               size_t workspace_bytes, aopt_stream_t stream) {
    if (n < 1 || b < 1 || !(grid_size > 0.f) || max_passes < 1 || max_passes > 6) return AOPT_ERR_INVALID_ARGUMENT;
    if (!coord || !offset || !order32 || !cluster32 || !cluster64 || !idx_ptr || !new_offset || !meta)
        return AOPT_ERR_INVALID_ARGUMENT;
    if (!workspace || workspace_bytes < aopt_voxel_grid_workspace_bytes(n, b)) return AOPT_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    char *ws = static_cast<char *>(workspace);
    unsigned *lo = reinterpret_cast<unsigned *>(ws); ws += a256(12 * (size_t)b);
    unsigned *hi = reinterpret_cast<unsigned *>(ws); ws += a256(12 * (size_t)b);
    float *start_f = reinterpret_cast<float *>(ws); ws += a256(12 * (size_t)b);
    unsigned long long *k0 = reinterpret_cast<unsigned long long *>(ws); ws += a256(8 * (size_t)n);
    unsigned long long *kp = reinterpret_cast<unsigned long long *>(ws); ws += a256(8 * (size_t)n);
    unsigned long long *kq = reinterpret_cast<unsigned long long *>(ws); ws += a256(8 * (size_t)n);
    int *vp = reinterpret_cast<int *>(ws); ws += a256(4 * (size_t)n);
    int *vq = reinterpret_cast<int *>(ws); ws += a256(4 * (size_t)n);
    int *flag = reinterpret_cast<int *>(ws); ws += a256(4 * ((size_t)n + 1));
    int *scan = reinterpret_cast<int *>(ws); ws += a256(4 * ((size_t)n + 1));
    int *scratch = reinterpret_cast<int *>(ws); ws += a256(4 * radix_scratch_ints(n) + 16);
    int *partial = reinterpret_cast<int *>(ws);

    // bbox -> layout -> keys -> (histogram, scan, scatter) x passes -> mark -> scan -> finalize: one chain of programmatic
    // dependent launches (tuning "pdl"; every kernel of the chain starts with pdl_wait())
    const bool pdl = tuning(kTunePdl) != 2;
    launch_scene_bbox(n, b, coord, offset, lo, hi, st);
    launch_chain(pdl, voxel_layout_kernel, 1, 128, 0, st, b, (const unsigned *)lo, (const unsigned *)hi, start, grid_size,
                 max_passes, start_f, meta);
    launch_chain(pdl, voxel_ckeys_kernel, div_up(n, kPoolBlock), kPoolBlock, 0, st, n, b, coord, offset, (const float *)start_f,
                 grid_size, k0, meta);
    const int *npass_dev = meta + 2;
    for (int pass = 0; pass < max_passes; ++pass) {
        unsigned long long *kout = (pass & 1) ? kq : kp;
        int *vout = (pass & 1) ? vq : vp;
        if (pass == 0)
            launch_radix_pass<unsigned long long>(PtrKeys<unsigned long long>{k0}, nullptr, kout, vout, n, pass, npass_dev, scratch, st, pdl);
        else
            launch_radix_pass<unsigned long long>(PtrKeys<unsigned long long>{(pass & 1) ? kp : kq}, (pass & 1) ? vp : vq, kout, vout,
                                                  n, pass, npass_dev, scratch, st, pdl);
    }
    const int grid = div_up(n > b ? n : b, kPoolBlock);
    launch_chain(pdl, voxel_mark2_kernel, div_up(n, kPoolBlock), kPoolBlock, 0, st, n, (const unsigned long long *)kp,
                 (const unsigned long long *)kq, (const int *)meta, flag, partial, (int)scan_partial_ints(n));
    launch_exclusive_scan_chained(flag, scan, n, partial, st, pdl);
    launch_chain(pdl, voxel_finalize2_kernel, grid, kPoolBlock, 0, st, n, b, (const int *)vp, (const int *)vq, (const int *)flag,
                 (const int *)scan, offset, order32, cluster32, cluster64, idx_ptr, new_offset, meta);
    return check_launch(6 + max_passes * kRadixLaunchesPerPass);
}

extern "C" int aopt_pool_forward(int n_vox, int c, const float *feat, const float *coord, const int *order,
                                 const int *idx_ptr, float *out_feat, int *argmax, float *out_coord,
                                 aopt_stream_t stream) {
    if (n_vox < 0 || c < 1) return AOPT_ERR_INVALID_ARGUMENT;
    if (n_vox == 0) return AOPT_OK;
    if (!feat || !order || !idx_ptr || !out_feat || !argmax || (out_coord && !coord)) return AOPT_ERR_INVALID_ARGUMENT;
    const bool vec = (c % 4 == 0) && aligned16(feat) && aligned16(out_feat) && aligned16(argmax);
    if (vec) {
        const int chunks = c / 4;
        pool_forward_kernel<4><<<col_grid(n_vox, chunks, kPoolBlock, 6), kPoolBlock, 0, as_stream(stream)>>>(
            n_vox, chunks, c, feat, coord, order, idx_ptr, out_feat, argmax, out_coord);
    } else {
        pool_forward_kernel<1><<<col_grid(n_vox, c, kPoolBlock, 8), kPoolBlock, 0, as_stream(stream)>>>(
            n_vox, c, c, feat, coord, order, idx_ptr, out_feat, argmax, out_coord);
    }
    return check_launch();
}

extern "C" int aopt_pool_backward(int n, int c, const float *grad_out, const int *argmax, const int *cluster,
                                  float *grad_feat, aopt_stream_t stream) {
    if (n < 0 || c < 1) return AOPT_ERR_INVALID_ARGUMENT;
    if (n == 0) return AOPT_OK;
    if (!grad_out || !argmax || !cluster || !grad_feat) return AOPT_ERR_INVALID_ARGUMENT;
    const bool vec = (c % 4 == 0) && aligned16(grad_out) && aligned16(argmax) && aligned16(grad_feat);
    if (vec) {
        const int chunks = c / 4;
        pool_backward_kernel<4><<<col_grid(n, chunks, kPoolBlock, 8), kPoolBlock, 0, as_stream(stream)>>>(
            n, chunks, c, grad_out, argmax, cluster, grad_feat);
    } else {
        pool_backward_kernel<1><<<col_grid(n, c, kPoolBlock, 8), kPoolBlock, 0, as_stream(stream)>>>(
            n, c, c, grad_out, argmax, cluster, grad_feat);
    }
    return check_launch();
}
