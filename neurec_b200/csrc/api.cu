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

extern "C" int nrc_stage_batch_host(const void* a_host, const void* b_host, const void* c_host,
                                    int64_t batch, void* staging, void* stream) {
    NRC_REQUIRE(batch > 0 && staging != nullptr, NRC_E_VALUE, "batch must be positive");
    cudaStream_t st = nrc::as_stream(stream);
    char* dst = reinterpret_cast<char*>(staging);
    const size_t nb = (size_t)batch * 4;
    const void* src[3] = {a_host, b_host, c_host};
    for (int i = 0; i < 3; ++i)
        if (src[i]) NRC_CUDA_CHECK(cudaMemcpyAsync(dst + i * nb, src[i], nb, cudaMemcpyHostToDevice, st));
    return NRC_OK;
}

extern "C" int nrc_fetch_host(const float* src_dev, float* dst_host, int64_t count, void* stream) {
    cudaStream_t st = nrc::as_stream(stream);
    NRC_CUDA_CHECK(cudaMemcpyAsync(dst_host, src_dev, (size_t)count * sizeof(float),
                                   cudaMemcpyDeviceToHost, st));
    NRC_CUDA_CHECK(cudaStreamSynchronize(st));
    return NRC_OK;
}

extern "C" int nrc_version(void) { return 100; }  // 0.1.0

extern "C" const char* nrc_last_error(void) { return nrc::g_err; }
