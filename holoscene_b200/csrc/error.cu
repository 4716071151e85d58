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

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return HSB_OK;
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return HSB_ERR_CUDA;
}
}  // namespace hsb

extern "C" const char* hsb_last_error(void) { return hsb::g_err; }
extern "C" int hsb_abi_version(void) { return HSB_ABI_VERSION; }
