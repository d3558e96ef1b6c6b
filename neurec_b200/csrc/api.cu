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

// ---- captured training-step graphs ------------------------------------------------------
struct nrc_step_graph {
    cudaGraph_t graph;
    cudaGraphExec_t exec;
};

extern "C" int nrc_graph_capture_begin(void* stream) {
    NRC_CUDA_CHECK(cudaStreamBeginCapture(nrc::as_stream(stream), cudaStreamCaptureModeThreadLocal));
    return NRC_OK;
}

extern "C" int nrc_graph_capture_end(void* stream, nrc_step_graph** out) {
    NRC_REQUIRE(out != nullptr, NRC_E_VALUE, "out is NULL");
    cudaGraph_t g = nullptr;
    NRC_CUDA_CHECK(cudaStreamEndCapture(nrc::as_stream(stream), &g));
    cudaGraphExec_t e = nullptr;
    cudaError_t err = cudaGraphInstantiate(&e, g, 0);
    if (err != cudaSuccess) {
        cudaGraphDestroy(g);
        nrc::set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(err));
        return NRC_E_CUDA;
    }
    *out = new nrc_step_graph{g, e};
    return NRC_OK;
}

extern "C" int nrc_graph_stage_async(const void* pinned_host, void* staging_dev, int64_t nbytes,
                                     void* stream) {
    NRC_CUDA_CHECK(cudaMemcpyAsync(staging_dev, pinned_host, (size_t)nbytes, cudaMemcpyHostToDevice,
                                   nrc::as_stream(stream)));
    return NRC_OK;
}

extern "C" int nrc_graph_fetch_async(const float* src_dev, float* pinned_host, int64_t count,
                                     void* stream) {
    NRC_CUDA_CHECK(cudaMemcpyAsync(pinned_host, src_dev, (size_t)count * sizeof(float),
                                   cudaMemcpyDeviceToHost, nrc::as_stream(stream)));
    return NRC_OK;
}

extern "C" int nrc_graph_step(nrc_step_graph* g, const void* a_host, const void* b_host,
                              const void* c_host, int64_t batch, float lr_t, void* pinned_stage,
                              void* stream) {
    NRC_REQUIRE(g != nullptr && pinned_stage != nullptr, NRC_E_VALUE, "graph / staging is NULL");
    char* dst = reinterpret_cast<char*>(pinned_stage);
    const size_t nb = (size_t)batch * 4;
    if (a_host) memcpy(dst, a_host, nb);
    if (b_host) memcpy(dst + nb, b_host, nb);
    if (c_host) memcpy(dst + 2 * nb, c_host, nb);
    memcpy(dst + 3 * nb, &lr_t, sizeof(float));
    cudaStream_t st = nrc::as_stream(stream);
    NRC_CUDA_CHECK(cudaGraphLaunch(g->exec, st));
    NRC_CUDA_CHECK(cudaStreamSynchronize(st));
    return NRC_OK;
}

extern "C" int nrc_graph_destroy(nrc_step_graph* g) {
    if (g) {
        cudaGraphExecDestroy(g->exec);
        cudaGraphDestroy(g->graph);
        delete g;
    }
    return NRC_OK;
}

extern "C" int nrc_version(void) { return 100; }  // 0.1.0

extern "C" const char* nrc_last_error(void) { return nrc::g_err; }
