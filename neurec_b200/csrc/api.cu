// Library-level entry points and shared host helpers.
#include <stdarg.h>

#include "common.cuh"

namespace nrc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;  // B200
    }
    return cached;
}

}  // namespace nrc

extern "C" int nrc_version(void) { return 100; }  // 0.1.0

extern "C" const char* nrc_last_error(void) { return nrc::g_err; }
