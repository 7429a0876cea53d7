// api.cu — status plumbing of the C ABI (include/ao_pointops.h).
#include <atomic>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace aopt {

static thread_local cudaError_t g_last_error = cudaSuccess;
static std::atomic<unsigned long long> g_kernel_launches{0};

int check_launch(int kernels) {
    g_kernel_launches.fetch_add((unsigned long long)kernels, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_last_error = e;
        return AOPT_ERR_LAUNCH;
    }
    return AOPT_OK;
}

static std::atomic<int> g_tuning[kTuneCount];
static std::atomic<bool> g_tuning_init{false};
static const char *const kTuneNames[kTuneCount] = {"csr_impl", "gva_bwd", "voxel_sort", "knn_topk", "knn_pend", "knn_sample", "pdl", "l2pf", "knn_site"};

static void tuning_init() {
    if (g_tuning_init.exchange(true)) return;
    auto env = [](const char *name, const char *one, const char *two) {
        const char *e = getenv(name);
        if (!e || !e[0]) return 0;
        return e[0] == one[0] ? 1 : e[0] == two[0] ? 2 : 0;
    };
    g_tuning[kTuneCsrImpl] = env("AOPT_CSR_IMPL", "sort", "count");
    g_tuning[kTuneGvaBwd] = env("AOPT_GVA_BWD", "fused", "split");
    g_tuning[kTuneVoxelSort] = env("AOPT_VOXEL_SORT", "radix", "wide");
    g_tuning[kTuneKnnTopk] = env("AOPT_KNN_TOPK", "heap", "list");
    g_tuning[kTuneKnnPend] = env("AOPT_KNN_PEND", "1", "0");
    g_tuning[kTuneKnnSample] = env("AOPT_KNN_SAMPLE", "bbox", "sampled");
    g_tuning[kTunePdl] = env("AOPT_PDL", "1", "0");
    g_tuning[kTuneL2Prefetch] = env("AOPT_L2PF", "1", "0");
    g_tuning[kTuneKnnSite] = env("AOPT_KNN_SITE", "1", "0");
}

int tuning(int which) {
    tuning_init();
    return (which >= 0 && which < kTuneCount) ? g_tuning[which].load(std::memory_order_relaxed) : 0;
}

}  // namespace aopt

extern "C" int aopt_set_tuning(const char *name, int value) {
    aopt::tuning_init();
    if (!name) return AOPT_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < aopt::kTuneCount; ++i) {
        if (aopt::kTuneNames[i] && strcmp(aopt::kTuneNames[i], name) == 0) {
            aopt::g_tuning[i].store(value, std::memory_order_relaxed);
            return AOPT_OK;
        }
    }
    return AOPT_ERR_INVALID_ARGUMENT;
}

extern "C" unsigned long long aopt_kernel_launches(void) {
    return aopt::g_kernel_launches.load(std::memory_order_relaxed);
}

extern "C" const char *aopt_version(void) { return "ao_pointops 0.1 (sm_100a)"; }

extern "C" const char *aopt_status_string(int status) {
    switch (status) {
        case AOPT_OK: return "ok";
        case AOPT_ERR_INVALID_ARGUMENT: return "invalid argument";
        case AOPT_ERR_WORKSPACE: return "workspace missing or too small";
        case AOPT_ERR_LAUNCH: return "CUDA launch error";
        case AOPT_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown status";
    }
}

extern "C" const char *aopt_last_cuda_error(void) { return cudaGetErrorString(aopt::g_last_error); }
