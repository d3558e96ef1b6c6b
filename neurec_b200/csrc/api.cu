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

// `to_stream` will not run past this point before everything issued so far on `from_stream` has
// completed.  Inside a stream capture this forks / joins the captured graph (the second stream
// joins the capture), which is how a burst graph overlaps the H2D of step s+1 with step s.
extern "C" int nrc_graph_depend(void* from_stream, void* to_stream) {
    static cudaEvent_t pool[512];
    static int next = 0;
    cudaEvent_t& ev = pool[next];
    next = (next + 1) % 512;
    if (!ev) NRC_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    NRC_CUDA_CHECK(cudaEventRecord(ev, nrc::as_stream(from_stream)));
    NRC_CUDA_CHECK(cudaStreamWaitEvent(nrc::as_stream(to_stream), ev, 0));
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

// A run of steps over consecutive batches of host arrays with a ring of `ring` pinned blocks.
// `burst` (optional) is ONE captured graph holding `ring` consecutive steps whose H2D / loss-D2H
// nodes sit on a second captured stream, so inside a burst the copy of step s+1 overlaps the
// kernels of step s; the host stages `ring` batches, launches it and waits.  Steps that do not fill
// a burst (and every step when burst == NULL) go through the per-slot single-step graphs, launched
// back to back with one wait per ring wrap.  Every step does its own H2D and its own loss D2H.
extern "C" int nrc_graph_run_steps(nrc_step_graph* burst, nrc_step_graph* const* graphs, int32_t ring,
                                   const void* a_host, const void* b_host, const void* c_host, int64_t batch,
                                   const float* lr_t, int64_t n_steps, void* const* pinned_stage,
                                   const float* const* loss_pinned, int32_t loss_count, double* loss_sum,
                                   void* stream) {
    NRC_REQUIRE(pinned_stage != nullptr && ring > 0, NRC_E_VALUE, "pinned ring is empty");
    NRC_REQUIRE(burst != nullptr || graphs != nullptr, NRC_E_VALUE, "no graph given");
    NRC_REQUIRE(lr_t != nullptr && n_steps >= 0 && batch > 0, NRC_E_VALUE, "bad step arguments");
    cudaStream_t st = nrc::as_stream(stream);
    const size_t nb = (size_t)batch * 4;
    double acc = 0.0;
    auto drain = [&](int64_t upto) {   // losses of the steps launched since the last wait
        for (int64_t r = 0; r < upto; ++r)
            if (loss_pinned && loss_pinned[r])
                for (int32_t k = 0; k < loss_count; ++k) acc += (double)loss_pinned[r][k];
    };
    auto stage = [&](int64_t s, int64_t r) {
        char* dst = reinterpret_cast<char*>(pinned_stage[r]);
        const size_t off = (size_t)s * nb;
        if (a_host) memcpy(dst, reinterpret_cast<const char*>(a_host) + off, nb);
        if (b_host) memcpy(dst + nb, reinterpret_cast<const char*>(b_host) + off, nb);
        if (c_host) memcpy(dst + 2 * nb, reinterpret_cast<const char*>(c_host) + off, nb);
        memcpy(dst + 3 * nb, lr_t + s, sizeof(float));
    };
    for (int64_t r = 0; r < ring; ++r)
        NRC_REQUIRE(pinned_stage[r] != nullptr, NRC_E_VALUE, "pinned ring slot %lld is NULL", (long long)r);
    int64_t s = 0;
    if (burst) {
        for (; s + ring <= n_steps; s += ring) {
            for (int64_t r = 0; r < ring; ++r) stage(s + r, r);
            NRC_CUDA_CHECK(cudaGraphLaunch(burst->exec, st));
            NRC_CUDA_CHECK(cudaStreamSynchronize(st));
            drain(ring);
        }
    }
    if (s < n_steps) {
        NRC_REQUIRE(graphs != nullptr, NRC_E_VALUE, "%lld steps do not fill a burst and no single-step graphs were given",
                    (long long)(n_steps - s));
        int64_t in_flight = 0;
        for (; s < n_steps; ++s) {
            const int64_t r = in_flight;
            NRC_REQUIRE(graphs[r] != nullptr, NRC_E_VALUE, "graph ring slot %lld is NULL", (long long)r);
            stage(s, r);
            NRC_CUDA_CHECK(cudaGraphLaunch(graphs[r]->exec, st));
            if (++in_flight == ring) {
                NRC_CUDA_CHECK(cudaStreamSynchronize(st));
                drain(in_flight);
                in_flight = 0;
            }
        }
        NRC_CUDA_CHECK(cudaStreamSynchronize(st));
        drain(in_flight);
    }
    if (loss_sum) *loss_sum = acc;
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

// ---- CUDA IPC for row-sharded tables -----------------------------------------------------------
// The exporter hands out the handle of the ALLOCATION that holds `dev_ptr` plus the offset of
// dev_ptr inside it (framework allocators sub-allocate).  The importer opens it with ITS OWN device
// current and cudaIpcMemLazyEnablePeerAccess, the way NCCL maps peer buffers, so that kernels of
// the importing device can dereference the mapping over NVLink.
extern "C" int nrc_ipc_export(const void* dev_ptr, void* handle64_out, int64_t* offset_out) {
    NRC_REQUIRE(dev_ptr && handle64_out && offset_out, NRC_E_VALUE, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);   // cuMemGetAddressRange_v2
    static range_fn get_range = nullptr;
    if (!get_range) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        NRC_CUDA_CHECK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q));
        NRC_REQUIRE(fn != nullptr && q == cudaDriverEntryPointSuccess, NRC_E_CUDA,
                    "the driver does not export cuMemGetAddressRange");
        get_range = reinterpret_cast<range_fn>(fn);
    }
    unsigned long long base = 0;
    size_t size = 0;
    const int cr = get_range(&base, &size, (unsigned long long)(uintptr_t)dev_ptr);
    NRC_REQUIRE(cr == 0 && base != 0, NRC_E_CUDA, "cuMemGetAddressRange failed (%d)", cr);
    cudaIpcMemHandle_t h;
    NRC_CUDA_CHECK(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>((uintptr_t)base)));
    memcpy(handle64_out, &h, sizeof(h));
    *offset_out = (int64_t)((unsigned long long)(uintptr_t)dev_ptr - base);
    return NRC_OK;
}

extern "C" int nrc_ipc_open(const void* handle64, int64_t offset, void** dev_ptr_out) {
    NRC_REQUIRE(handle64 && dev_ptr_out && offset >= 0, NRC_E_VALUE, "bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* base = nullptr;
    NRC_CUDA_CHECK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr_out = reinterpret_cast<char*>(base) + offset;
    return NRC_OK;
}

extern "C" int nrc_ipc_close(void* dev_ptr, int64_t offset) {
    if (!dev_ptr) return NRC_OK;
    NRC_CUDA_CHECK(cudaIpcCloseMemHandle(reinterpret_cast<char*>(dev_ptr) - offset));
    return NRC_OK;
}

extern "C" int nrc_shard_alloc(int64_t nbytes, void** dev_ptr_out, void* handle64_out) {
    NRC_REQUIRE(nbytes > 0 && dev_ptr_out && handle64_out, NRC_E_VALUE, "bad argument");
    void* p = nullptr;
    NRC_CUDA_CHECK(cudaMalloc(&p, (size_t)nbytes));
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        nrc::set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return NRC_E_CUDA;
    }
    memcpy(handle64_out, &h, sizeof(h));
    *dev_ptr_out = p;
    return NRC_OK;
}

extern "C" int nrc_shard_free(void* dev_ptr) {
    if (dev_ptr) NRC_CUDA_CHECK(cudaFree(dev_ptr));
    return NRC_OK;
}

extern "C" int nrc_version(void) { return 100; }  // 0.1.0

extern "C" const char* nrc_last_error(void) { return nrc::g_err; }
