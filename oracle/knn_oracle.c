/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the product path
 * (ao_b200/).  Allowed users: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
 * --impl reference legs.
 *
 * CPU restatement (plain C) of the reference's batched-offset kNN:
 *   /root/reference/libs/pointops/src/knn_query/knn_query_cuda_kernel.cu:60-104
 *
 * Pinning: the reference has no golden vectors for this path.  The oracle is
 * pinned against the reference itself: oracle/Makefile.ref compiles the UNMODIFIED reference
 * launchers into oracle/_ref/libpointops_ref.so, and tests/test_knn_gpu.py compares this file
 * with that library on the B200 (bit-exact idx and dist2 on tie-free rows); the outputs of that
 * run are committed as fixtures under tests/golden/ (see tests/golden/README.md).
 *
 * Arithmetic.  The reference source line (:92) is
 *     d2 = (nx-x)*(nx-x) + (ny-y)*(ny-y) + (nz-z)*(nz-z)
 * and nvcc's default -fmad=true contracts it.  SASS of the reference kernel built here
 * (nvcc 12.9.86, -O2, sm_100a):  FADD dy; FADD dx; FMUL dy*dy; FADD dz; FFMA dx*dx+.; FFMA dz*dz+.
 * i.e.  d2 = fma(dz, dz, fma(dx, dx, dy*dy)),  with d = query - candidate.
 * (SURVEY.md §7/§8 lists the x and y terms the other way round; the compiled reference wins.)
 *
 * Two selection rules are provided:
 *   knn_oracle_heap : the reference's exact algorithm — k-slot max-heap, strict '<' accept
 *                     (:93), reheap (:15-30), heap_sort ascending (:33-42), init 1e10 / -1
 *                     (:85-86).  On exact d2 ties the kept set/order is heap-state dependent.
 *   knn_oracle_lex  : the new deterministic contract of BASELINE.json's north_star — ascending
 *                     (d2, idx) lexicographic ("ties broken by lower index").  Identical to
 *                     knn_oracle_heap on every row with no equal d2 among its k+1 smallest.
 * Segment lookup follows :45-56 and :70-76 (offset = cumulative END indices).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#define ORACLE_MAX_K 128 /* knn_query_cuda_kernel.cu:82-83: float best_dist[128] */

static inline float dist2_ref(float qx, float qy, float qz, float x, float y, float z) {
    float dx = qx - x, dy = qy - y, dz = qz - z;
    float t = dy * dy;          /* FMUL */
    t = fmaf(dx, dx, t);        /* FFMA */
    return fmaf(dz, dz, t);     /* FFMA */
}

/* knn_query_cuda_kernel.cu:45-56 */
static int get_bt_idx(int idx, const int *offset) {
    int i = 0;
    while (1) {
        if (idx < offset[i]) break;
        else i++;
    }
    return i;
}

/* knn_query_cuda_kernel.cu:15-30 */
static void reheap(float *dist, int *idx, int k) {
    int root = 0;
    int child = root * 2 + 1;
    while (child < k) {
        if (child + 1 < k && dist[child + 1] > dist[child]) child++;
        if (dist[root] > dist[child]) return;
        float td = dist[root]; dist[root] = dist[child]; dist[child] = td;
        int ti = idx[root]; idx[root] = idx[child]; idx[child] = ti;
        root = child;
        child = root * 2 + 1;
    }
}

/* knn_query_cuda_kernel.cu:33-42 */
static void heap_sort(float *dist, int *idx, int k) {
    for (int i = k - 1; i > 0; i--) {
        float td = dist[0]; dist[0] = dist[i]; dist[i] = td;
        int ti = idx[0]; idx[0] = idx[i]; idx[i] = ti;
        reheap(dist, idx, i);
    }
}

/* Reference algorithm, one query.  Returns 0, or -1 if nsample is out of range. */
int knn_oracle_heap(int m, int nsample, const float *xyz, const float *new_xyz,
                    const int *offset, const int *new_offset, int *idx, float *dist2) {
    if (nsample < 1 || nsample > ORACLE_MAX_K) return -1;
    for (int pt = 0; pt < m; pt++) {
        int bt = get_bt_idx(pt, new_offset);
        int start = bt == 0 ? 0 : offset[bt - 1];
        int end = offset[bt];
        float qx = new_xyz[pt * 3 + 0], qy = new_xyz[pt * 3 + 1], qz = new_xyz[pt * 3 + 2];
        float best_dist[ORACLE_MAX_K];
        int best_idx[ORACLE_MAX_K];
        for (int i = 0; i < nsample; i++) { best_dist[i] = 1e10f; best_idx[i] = -1; }
        for (int i = start; i < end; i++) {
            float d2 = dist2_ref(qx, qy, qz, xyz[i * 3 + 0], xyz[i * 3 + 1], xyz[i * 3 + 2]);
            if (d2 < best_dist[0]) {
                best_dist[0] = d2;
                best_idx[0] = i;
                reheap(best_dist, best_idx, nsample);
            }
        }
        heap_sort(best_dist, best_idx, nsample);
        for (int i = 0; i < nsample; i++) {
            idx[(size_t)pt * nsample + i] = best_idx[i];
            dist2[(size_t)pt * nsample + i] = best_dist[i];
        }
    }
    return 0;
}

/* Deterministic contract: ascending (d2, idx).  Same distance formula, same 1e10 / -1 padding,
 * same strict accept against the running k-th distance. */
int knn_oracle_lex(int m, int nsample, const float *xyz, const float *new_xyz,
                   const int *offset, const int *new_offset, int *idx, float *dist2) {
    if (nsample < 1 || nsample > ORACLE_MAX_K) return -1;
    for (int pt = 0; pt < m; pt++) {
        int bt = get_bt_idx(pt, new_offset);
        int start = bt == 0 ? 0 : offset[bt - 1];
        int end = offset[bt];
        float qx = new_xyz[pt * 3 + 0], qy = new_xyz[pt * 3 + 1], qz = new_xyz[pt * 3 + 2];
        float bd[ORACLE_MAX_K];
        int bi[ORACLE_MAX_K];
        for (int i = 0; i < nsample; i++) { bd[i] = 1e10f; bi[i] = -1; }
        for (int i = start; i < end; i++) {
            float d2 = dist2_ref(qx, qy, qz, xyz[i * 3 + 0], xyz[i * 3 + 1], xyz[i * 3 + 2]);
            if (d2 < bd[nsample - 1]) { /* candidates arrive in ascending idx: equal d2 never displaces */
                int j = nsample - 1;
                while (j > 0 && bd[j - 1] > d2) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; j--; }
                bd[j] = d2;
                bi[j] = i;
            }
        }
        for (int i = 0; i < nsample; i++) {
            idx[(size_t)pt * nsample + i] = bi[i];
            dist2[(size_t)pt * nsample + i] = bd[i];
        }
    }
    return 0;
}

/* Same as knn_oracle_lex for queries [q_begin, q_end) only — lets callers spread the work over
 * host threads (bench.py cpu_baseline) or check a sample of rows at full problem size. */
int knn_oracle_lex_range(int q_begin, int q_end, int nsample, const float *xyz,
                         const float *new_xyz, const int *offset, const int *new_offset,
                         int *idx, float *dist2) {
    if (nsample < 1 || nsample > ORACLE_MAX_K) return -1;
    for (int pt = q_begin; pt < q_end; pt++) {
        int bt = get_bt_idx(pt, new_offset);
        int start = bt == 0 ? 0 : offset[bt - 1];
        int end = offset[bt];
        float qx = new_xyz[pt * 3 + 0], qy = new_xyz[pt * 3 + 1], qz = new_xyz[pt * 3 + 2];
        float bd[ORACLE_MAX_K];
        int bi[ORACLE_MAX_K];
        for (int i = 0; i < nsample; i++) { bd[i] = 1e10f; bi[i] = -1; }
        for (int i = start; i < end; i++) {
            float d2 = dist2_ref(qx, qy, qz, xyz[i * 3 + 0], xyz[i * 3 + 1], xyz[i * 3 + 2]);
            if (d2 < bd[nsample - 1]) {
                int j = nsample - 1;
                while (j > 0 && bd[j - 1] > d2) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; j--; }
                bd[j] = d2;
                bi[j] = i;
            }
        }
        for (int i = 0; i < nsample; i++) {
            idx[(size_t)(pt - q_begin) * nsample + i] = bi[i];
            dist2[(size_t)(pt - q_begin) * nsample + i] = bd[i];
        }
    }
    return 0;
}
