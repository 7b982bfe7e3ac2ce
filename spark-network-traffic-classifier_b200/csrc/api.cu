// api.cu — error reporting shared by every entry point of libb200flow.so.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace b200flow {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return B200FLOW_ERR_CUDA;
    }
    return B200FLOW_OK;
}

}  // namespace b200flow

extern "C" const char* b200flow_last_error(void) { return b200flow::g_err; }
extern "C" int b200flow_version(void) { return 100; }
