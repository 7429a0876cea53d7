// datapipe.cu — the per-sample data transforms that feed the PTv2m2 hot path, on the GPU:
//   GridSample  /root/reference/pointcept/datasets/transform.py:792-896  (voxel hash, argsort, unique, pick)
//   SphereCrop  /root/reference/pointcept/datasets/transform.py:968-979  (squared distance to a centre, argsort)
// The reference runs them in numpy inside 16 dataloader workers per GPU (configs/_base_/default_runtime.py:8);
// a raw S3DIS room is 0.5-1 M points, so one sample costs ~0.3-0.5 s of CPU time against a ~10 ms GPU step.
//
// Kernels (integer / byte work, HBM-bound, one thread per point, coalesced 12-byte rows):
//   grid_discretize_kernel   floor(coord / grid) per axis (fp64 or fp32 division, see `f64`) + per-axis minimum
//   grid_hash_kernel         subtract the minimum, FNV64-1A (or ravel) key per point; key ^ 2^63 so that an
//                            int64 radix sort orders like the reference's uint64 argsort
//   sphere_dist2_kernel      (x-cx)^2 + (y-cy)^2 + (z-cz)^2 with numpy's operation order (no contraction)
//   select_rows_kernel       out[i, :] = src[index[i], :]  (gathers every per-point key of the sample)
// The 64-bit / 32-bit key sorts are the library radix sort (torch.sort, stable), and the voxel partition after
// the sort is aopt_voxel_partition (pool.cu) with a single scene.
//
// Division semantics.  transform.py:806 is `data_dict["coord"] / np.array(self.grid_size)`: fp32 array by a 0-d
// fp64 array.  NumPy >= 2 (NEP 50; what this image runs, and what generated tests/golden/datapipe_*.npz)
// promotes to fp64; NumPy 1.x value-based casting kept fp32.  f64 = 1 / 0 selects either, bit-exactly.
#include "common.cuh"

namespace aopt {

constexpr int kDpBlock = 256;

template <bool F64>
__global__ void __launch_bounds__(kDpBlock)
grid_discretize_kernel(int n, const float *__restrict__ coord, double gx, double gy, double gz,
                       int *__restrict__ cell, int *__restrict__ cell_min) {
    __shared__ int red[3][kDpBlock / 32];
    const int i = blockIdx.x * kDpBlock + threadIdx.x;
    int c[3] = {INT_MAX, INT_MAX, INT_MAX};
    if (i < n) {
        const double g[3] = {gx, gy, gz};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = __ldg(coord + (size_t)i * 3 + a);
            // np.floor(scaled).astype(int): floor in the division's dtype, then a C cast
            c[a] = F64 ? (int)floor(__ddiv_rn((double)v, g[a])) : (int)floorf(__fdiv_rn(v, (float)g[a]));
            cell[(size_t)i * 3 + a] = c[a];
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        int m = c[a];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, d));
        if ((threadIdx.x & 31) == 0) red[a][threadIdx.x >> 5] = m;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        int m = red[threadIdx.x][0];
        for (int w = 1; w < kDpBlock / 32; ++w) m = min(m, red[threadIdx.x][w]);
        atomicMin(cell_min + threadIdx.x, m);
    }
}

// hash_type 0: FNV64-1A as written at transform.py:881-896 (multiply by the prime, then xor, per axis);
// hash_type 1: ravel (transform.py:864-878): ((x * max_y1) + y) * max_z1 + z with max_*1 = per-axis max + 1.
__global__ void __launch_bounds__(kDpBlock)
grid_hash_kernel(int n, int *__restrict__ cell, const int *__restrict__ cell_min, int hash_type,
                 const int *__restrict__ cell_max, int64_t *__restrict__ keys) {
    const int i = blockIdx.x * kDpBlock + threadIdx.x;
    if (i >= n) return;
    unsigned long long v[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int c = cell[(size_t)i * 3 + a] - __ldg(cell_min + a);   // discrete_coord -= discrete_coord.min(0)
        cell[(size_t)i * 3 + a] = c;
        v[a] = (unsigned long long)(long long)c;                       // astype(np.uint64)
    }
    unsigned long long h;
    if (hash_type == 0) {
        h = 14695981039346656037ull;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            h *= 1099511628211ull;
            h ^= v[a];
        }
    } else {
        // arr -= arr.min(0) was already applied; arr_max = arr.max(0) + 1
        const unsigned long long my = (unsigned long long)(__ldg(cell_max + 1) - __ldg(cell_min + 1)) + 1ull;
        const unsigned long long mz = (unsigned long long)(__ldg(cell_max + 2) - __ldg(cell_min + 2)) + 1ull;
        h = (v[0] * my + v[1]) * mz + v[2];
    }
    keys[i] = (int64_t)(h ^ 0x8000000000000000ull);
}

__global__ void __launch_bounds__(kDpBlock)
cell_max_kernel(int n, const int *__restrict__ cell, int *__restrict__ cell_max) {
    const int i = blockIdx.x * kDpBlock + threadIdx.x;
    int c[3] = {INT_MIN, INT_MIN, INT_MIN};
    if (i < n) {
#pragma unroll
        for (int a = 0; a < 3; ++a) c[a] = cell[(size_t)i * 3 + a];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        int m = c[a];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
        if ((threadIdx.x & 31) == 0) atomicMax(cell_max + a, m);
    }
}

// np.sum(np.square(coord - center), 1) in fp32: three subtractions, three squares, ((sx + sy) + sz).
__global__ void __launch_bounds__(kDpBlock)
sphere_dist2_kernel(int n, const float *__restrict__ coord, float cx, float cy, float cz,
                    float *__restrict__ dist2) {
    const int i = blockIdx.x * kDpBlock + threadIdx.x;
    if (i >= n) return;
    const float dx = __fsub_rn(__ldg(coord + (size_t)i * 3), cx);
    const float dy = __fsub_rn(__ldg(coord + (size_t)i * 3 + 1), cy);
    const float dz = __fsub_rn(__ldg(coord + (size_t)i * 3 + 2), cz);
    dist2[i] = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// out[i, 0..w) = src[index[i], 0..w) for rows of w 4-byte words (any 32-bit dtype).
__global__ void __launch_bounds__(kDpBlock)
select_rows_kernel(long long total, int w, const uint32_t *__restrict__ src, const int64_t *__restrict__ index,
                   uint32_t *__restrict__ out) {
    const long long step = (long long)gridDim.x * kDpBlock;
    for (long long t = (long long)blockIdx.x * kDpBlock + threadIdx.x; t < total; t += step) {
        const long long r = t / w;
        const int col = (int)(t - r * w);
        out[t] = __ldg(src + (size_t)__ldg(index + r) * w + col);
    }
}

// pick[v] = order[idx_ptr[v] + r[v] % count[v]]   (transform.py:813-817, train mode), or with r[v] = part for
// test mode (:841-843).  order = stable argsort of the keys, idx_ptr = voxel boundaries.
__global__ void __launch_bounds__(kDpBlock)
voxel_pick_kernel(int n_vox, const int *__restrict__ idx_ptr, const int *__restrict__ order,
                  const int64_t *__restrict__ r, long long r_const, int64_t *__restrict__ pick) {
    const int v = blockIdx.x * kDpBlock + threadIdx.x;
    if (v >= n_vox) return;
    const int s = __ldg(idx_ptr + v), cnt = __ldg(idx_ptr + v + 1) - s;
    const long long rv = r ? __ldg(r + v) : r_const;
    pick[v] = (int64_t)__ldg(order + s + (int)(rv % cnt));
}

}  // namespace aopt

using namespace aopt;

// cell (n,3) int32 out = discrete coordinates minus their per-axis minimum (transform.py:807-809);
// keys (n) int64 out = hash ^ 2^63; stats (6) int32 out = per-axis min then max of floor(coord / grid)
// (min_coord = stats[0..2] * grid, transform.py:808).  grid = per-axis cell size.
extern "C" int aopt_grid_sample_keys(int n, const float *coord, double grid_x, double grid_y, double grid_z,
                                     int f64, int hash_type, int *cell, int64_t *keys, int *stats,
                                     aopt_stream_t stream) {
    if (n < 0 || !(grid_x > 0.0) || !(grid_y > 0.0) || !(grid_z > 0.0) || hash_type < 0 || hash_type > 1)
        return AOPT_ERR_INVALID_ARGUMENT;
    if (n == 0) return AOPT_OK;
    if (!coord || !cell || !keys || !stats) return AOPT_ERR_INVALID_ARGUMENT;
    cudaStream_t st = as_stream(stream);
    const int grid = div_up(n, kDpBlock);
    cudaMemsetAsync(stats, 0x7f, 3 * sizeof(int), st);        // 0x7f7f7f7f: above any cell index
    cudaMemsetAsync(stats + 3, 0x80, 3 * sizeof(int), st);    // 0x80808080: below any cell index
    if (f64) grid_discretize_kernel<true><<<grid, kDpBlock, 0, st>>>(n, coord, grid_x, grid_y, grid_z, cell, stats);
    else grid_discretize_kernel<false><<<grid, kDpBlock, 0, st>>>(n, coord, grid_x, grid_y, grid_z, cell, stats);
    cell_max_kernel<<<grid, kDpBlock, 0, st>>>(n, cell, stats + 3);
    grid_hash_kernel<<<grid, kDpBlock, 0, st>>>(n, cell, stats, hash_type, stats + 3, keys);
    return check_launch(3);
}

extern "C" int aopt_voxel_pick(int n_vox, const int *idx_ptr, const int *order, const int64_t *r,
                               long long r_const, int64_t *pick, aopt_stream_t stream) {
    if (n_vox < 0) return AOPT_ERR_INVALID_ARGUMENT;
    if (n_vox == 0) return AOPT_OK;
    if (!idx_ptr || !order || !pick || r_const < 0) return AOPT_ERR_INVALID_ARGUMENT;
    voxel_pick_kernel<<<div_up(n_vox, kDpBlock), kDpBlock, 0, as_stream(stream)>>>(n_vox, idx_ptr, order, r, r_const, pick);
    return check_launch();
}

extern "C" int aopt_sphere_dist2(int n, const float *coord, float cx, float cy, float cz, float *dist2,
                                 aopt_stream_t stream) {
    if (n < 0) return AOPT_ERR_INVALID_ARGUMENT;
    if (n == 0) return AOPT_OK;
    if (!coord || !dist2) return AOPT_ERR_INVALID_ARGUMENT;
    sphere_dist2_kernel<<<div_up(n, kDpBlock), kDpBlock, 0, as_stream(stream)>>>(n, coord, cx, cy, cz, dist2);
    return check_launch();
}

extern "C" int aopt_select_rows(long long rows, int words_per_row, const void *src, const int64_t *index,
                                void *out, aopt_stream_t stream) {
    if (rows < 0 || words_per_row < 1) return AOPT_ERR_INVALID_ARGUMENT;
    if (rows == 0) return AOPT_OK;
    if (!src || !index || !out) return AOPT_ERR_INVALID_ARGUMENT;
    const long long total = rows * words_per_row;
    select_rows_kernel<<<stride_grid(total, kDpBlock, 8), kDpBlock, 0, as_stream(stream)>>>(
        total, words_per_row, static_cast<const uint32_t *>(src), index, static_cast<uint32_t *>(out));
    return check_launch();
}
