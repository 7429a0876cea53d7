// api.cu — status plumbing of the C ABI (include/ao_pointops.h).
#include <atomic>

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

}  // namespace aopt

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
