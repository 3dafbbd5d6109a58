// abi.cu -- error string, version and device selection of the C ABI.
#include <stdarg.h>

#include "common.cuh"

namespace iv {
static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
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
extern "C" uint64_t iv_launch_count(void) { return __atomic_load_n(&iv::g_launches, __ATOMIC_RELAXED); }
