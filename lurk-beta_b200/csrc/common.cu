// Error reporting and device discovery for liblurk_b200 (see include/lurk_b200.h for conventions).
#include "common.cuh"

namespace lurk {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int require_gpu() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); liblurk_b200 has no CPU fallback", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
        return LURK_ERR_NOGPU;
    }
    return LURK_OK;
}

}  // namespace lurk

extern "C" {

const char *lurk_last_error(void) { return lurk::g_err; }
int lurk_version(void) { return 100; }
int lurk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int lurk_field_modulus(int field_id, uint8_t out[32]) {
    return lurk::dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        F m = F::modulus_raw();
        memcpy(out, m.v, 32);
        return LURK_OK;
    });
}

}  // extern "C"
