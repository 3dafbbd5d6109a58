// abi.cu -- error string, version and device selection of the C ABI.
#include <stdarg.h>

#include "common.cuh"

namespace iv {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace iv

extern "C" int iv_abi_version(void) { return IV_ABI_VERSION; }
extern "C" const char* iv_last_error_string(void) { return iv::g_err; }
extern "C" int iv_set_device(int device) {
    IV_CUDA(cudaSetDevice(device));
    return IV_OK;
}
