// Error reporting for the C ABI: int status + thread-local message (hsb_last_error).
#include "common.cuh"
#include "../../include/hsb200.h"
#include <stdio.h>
#include <string.h>

namespace hsb {
static thread_local char g_err[512] = "";

void set_error(const char* msg) {
    strncpy(g_err, msg, sizeof(g_err) - 1);
    g_err[sizeof(g_err) - 1] = 0;
}

static unsigned long long g_launches = 0;
void count_launch(int n) { g_launches += (unsigned long long)n; }

int check_cuda(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return HSB_OK;
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return HSB_ERR_CUDA;
}

int check_launch(const char* what) {
    g_launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return HSB_OK;
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return HSB_ERR_CUDA;
}
}  // namespace hsb

extern "C" const char* hsb_last_error(void) { return hsb::g_err; }
extern "C" int hsb_abi_version(void) { return HSB_ABI_VERSION; }
extern "C" unsigned long long hsb_launch_count(void) { return hsb::g_launches; }
