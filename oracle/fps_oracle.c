/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the product path
 * (ao_b200/).  Allowed users: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline legs.
 *
 * CPU restatement (plain C) of the reference's farthest point sampling:
 *   /root/reference/libs/pointops/src/sampling/sampling_cuda_kernel.cu:14-122  (kernel)
 *   /root/reference/libs/pointops/src/cuda_utils.h:11-14                        (block size)
 *   /root/reference/libs/pointops/functions/sampling.py:9-24                    (tmp = 1e10, n_max)
 *
 * The kernel is emulated thread by thread: `block` threads, thread tid scans k = start+tid,
 * start+tid+block, ... keeping the first strict maximum of min(d, tmp[k]) (:47-57, best = -1,
 * besti = start_n), then the shared-memory tree (:62-120) merges pairs (tid, tid+half) keeping the
 * lower tid unless the upper value is strictly greater (__update, :5-10).  Emulating the tree literally
 * (instead of restating "lowest (tid, k) wins") keeps the oracle independent of the argument the CUDA
 * implementation relies on.
 *
 * Arithmetic: source line :51 is (x2-x1)^2 + (y2-y1)^2 + (z2-z1)^2 contracted by nvcc's default
 * -fmad=true.  SASS of the reference kernel built here (nvcc 12.9.86, -O2, sm_100a):
 *   FADD dy; FADD dx; FMUL dy*dy; FADD dz; FFMA dx*dx+.; FFMA dz*dz+.; FMNMX with tmp.
 *
 * Pinning: tests/test_fps_gpu.py compares this file with the UNMODIFIED reference launcher
 * (oracle/_ref/libpointops_ref.so) on the B200, and tests/golden/fps_ref_cuda.npz holds outputs of that
 * launcher for seeded cases so that the CPU suite pins the oracle without a GPU.
 */
#include <math.h>
#include <stdlib.h>

static int ref_block_size(int n_max) { /* cuda_utils.h:11-14: min(2^floor(log2 n), 1024), at least 1 */
    int pow_2 = (int)(log((double)n_max) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

/* Returns 0 on success.  tmp (n) must be pre-filled by the caller (1e10), idx has new_offset[b-1] slots. */
int fps_oracle(int b, int n_max, const float *xyz, const int *offset, const int *new_offset, float *tmp,
               int *idx) {
    const int block = n_max > 0 ? ref_block_size(n_max) : 1;
    float *dists = (float *)malloc(sizeof(float) * (size_t)block);
    int *dists_i = (int *)malloc(sizeof(int) * (size_t)block);
    if (!dists || !dists_i) { free(dists); free(dists_i); return 1; }
    for (int bid = 0; bid < b; ++bid) {
        const int start_n = bid == 0 ? 0 : offset[bid - 1], end_n = offset[bid];
        const int start_m = bid == 0 ? 0 : new_offset[bid - 1], end_m = new_offset[bid];
        int old = start_n;
        if (end_m <= start_m) continue; /* the reference would write idx[start_m] out of the scene's range */
        idx[start_m] = start_n;
        for (int j = start_m + 1; j < end_m; ++j) {
            const float x1 = xyz[old * 3 + 0], y1 = xyz[old * 3 + 1], z1 = xyz[old * 3 + 2];
            for (int tid = 0; tid < block; ++tid) {
                int besti = start_n;
                float best = -1.f;
                for (int k = start_n + tid; k < end_n; k += block) {
                    const float dy = xyz[k * 3 + 1] - y1, dx = xyz[k * 3 + 0] - x1, dz = xyz[k * 3 + 2] - z1;
                    float d = dy * dy;
                    d = fmaf(dx, dx, d);
                    d = fmaf(dz, dz, d);
                    const float d2 = fminf(d, tmp[k]);
                    tmp[k] = d2;
                    besti = d2 > best ? k : besti;
                    best = d2 > best ? d2 : best;
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            for (int half = block / 2; half >= 1; half /= 2) {
                for (int tid = 0; tid < half; ++tid) {
                    const float v1 = dists[tid], v2 = dists[tid + half];
                    const int i1 = dists_i[tid], i2 = dists_i[tid + half];
                    dists[tid] = v1 > v2 ? v1 : v2;
                    dists_i[tid] = v2 > v1 ? i2 : i1;
                }
            }
            old = dists_i[0];
            idx[j] = old;
        }
    }
    free(dists);
    free(dists_i);
    return 0;
}
