// Host-side check of the register top-k insert used by the kNN kernels (ao_b200/csrc/knn_common.cuh).
// The struct is compiled for the HOST here (AOPT_TOPK_HD override) and fed random candidate streams with many
// exact distance ties; the result must equal a stable sort by (d2, idx) — the LEX rule — or, for LEX = false
// with candidates offered in ascending index, the same thing.  Built and run by tests/test_topk_host.py (no GPU).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define AOPT_TOPK_HD __host__ __device__ inline
static inline __host__ __device__ unsigned aopt_f2u_host(float x) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(x);
#else
    unsigned u;
    memcpy(&u, &x, 4);
    return u;
#endif
}
#define AOPT_F2U(x) aopt_f2u_host(x)
#include "../../ao_b200/csrc/knn_common.cuh"

using aopt::TopK;

struct Cand { float d; int i; };

template <int K, bool LEX>
static int run_case(unsigned seed, int n, int levels) {
    srand(seed);
    std::vector<Cand> c(n);
    for (int t = 0; t < n; ++t) {
        // few distinct distances -> many exact ties; a few values at / beyond the 1e10 sentinel and a NaN
        float d = (float)(rand() % levels) * 0.25f;
        int r = rand() % 97;
        if (r == 0) d = 1e10f;
        if (r == 1) d = 3e10f;
        if (r == 2) d = nanf("");
        c[t] = {d, t};
    }
    std::vector<Cand> order = c;
    if (LEX) {  // arbitrary arrival order
        for (int t = n - 1; t > 0; --t) std::swap(order[t], order[rand() % (t + 1)]);
    }
    TopK<K, LEX> top;
    top.init();
    for (const Cand &x : order) top.offer(x.d, x.i);
    // reference: candidates with d < 1e10 (NaN excluded), sorted by (d, i), first K, padded with (1e10, -1)
    std::vector<Cand> ok;
    for (const Cand &x : c) if (x.d < 1e10f) ok.push_back(x);
    std::stable_sort(ok.begin(), ok.end(), [](const Cand &a, const Cand &b) { return a.d < b.d || (a.d == b.d && a.i < b.i); });
    int bad = 0;
    for (int j = 0; j < K; ++j) {
        float rd = j < (int)ok.size() ? ok[j].d : 1e10f;
        int ri = j < (int)ok.size() ? ok[j].i : -1;
        if (!(top.d[j] == rd) || top.id[j] != ri) ++bad;
    }
    return bad;
}

template <int K>
static int run_k() {
    int bad = 0;
    for (unsigned s = 1; s <= 200; ++s) {
        const int n = (s % 7 == 0) ? (int)(s % (K + 3)) : 20 + (int)(s * 37 % 400);   // includes n < K and n = 0
        const int levels = 1 + (int)(s % 50);
        bad += run_case<K, true>(s, n, levels);
        bad += run_case<K, false>(s, n, levels);
    }
    return bad;
}

int main() {
    int bad = run_k<1>() + run_k<3>() + run_k<4>() + run_k<8>() + run_k<16>() + run_k<32>();
    printf("topk_host_test: %s (%d mismatching slots)\n", bad ? "FAIL" : "ok", bad);
    return bad ? 1 : 0;
}
