// Library-level C ABI: version, per-thread error text, device check.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace detr {
static thread_local char g_err[512] = {0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool pdl_enabled() {
    static const bool on = []() { const char* e = getenv("DETR_B200_NO_PDL"); return !(e && e[0] && e[0] != '0'); }();
    return on;
}
}  // namespace detr

extern "C" int detr_b200_abi_version(void) { return DETR_B200_ABI_VERSION; }

extern "C" int detr_b200_last_error(char* buf, int n) {
    if (!buf || n <= 0) return 0;
    strncpy(buf, detr::g_err, (size_t)n - 1);
    buf[n - 1] = 0;
    return (int)strlen(buf);
}

extern "C" int detr_b200_check_device(int device) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        detr::set_error("cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
        return 2;
    }
    if (prop.major != 10) {
        detr::set_error("device %d is sm_%d%d; libdetr_b200 contains sm_100a code only", device, prop.major, prop.minor);
        return 1;
    }
    return 0;
}
