// Peer-memory probe (measurement aid, not part of the library): what does a B200 sustain on RANDOM 512-byte rows of a
// table that lives on the NVLink peer -- by access method, by address range, by rows in flight?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/peer_probe profiles/peer_probe.cu && gpurun_out/peer_probe
// One process, two devices (cudaDeviceEnablePeerAccess): device 0 runs the kernels, the table sits on device 1
// ("peer") or on device 0 ("local", the control).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int D = 128;                       // floats per row (512 B)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ int64_t row_of(int64_t k, int64_t rows, int sequential) {
    if (sequential == 1) return k % rows;
    if (sequential == 2) return 12345 % rows;                  // one hot row
    if (sequential == 3) {                                     // p(rank) ~ 1/rank (Zipf exponent 1): rank = rows^u
        const float u = (mix((uint32_t)k * 2654435761u + 17u) >> 8) * (1.0f / 16777216.0f);
        const int64_t rank = (int64_t)exp2f(u * log2f((float)rows));
        return (rank < rows ? rank : rows - 1);
    }
    const uint64_t r = ((uint64_t)mix((uint32_t)k) << 20) ^ mix((uint32_t)(k >> 3) + 0x9e3779b9u);
    return (int64_t)(r % (uint64_t)rows);
}

// mode 0: warp per row, LDG.128 per lane, R rows in flight per warp
template <int R>
__global__ void __launch_bounds__(256) read_ldg(const float* __restrict__ T, int64_t rows, int64_t n, int seq, float* sink) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float acc = 0.f;
    for (int64_t k = warp * R; k < n; k += nw * R) {
        float4 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = __ldcg(reinterpret_cast<const float4*>(T + row_of(k + r, rows, seq) * D) + lane);
#pragma unroll
        for (int r = 0; r < R; ++r) acc += v[r].x + v[r].y + v[r].z + v[r].w;
    }
    if (acc == 12345.678f && sink) *sink = acc;
}

// mode 1: warp per row, RED.v4.f32 (no read)
__global__ void __launch_bounds__(256) red_v4(float* T, int64_t rows, int64_t n, int seq) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = warp; k < n; k += nw) {
        float* p = T + row_of(k, rows, seq) * D + lane * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1e-9f), "f"(1e-9f), "f"(1e-9f), "f"(1e-9f) : "memory");
    }
}

// scalar REDs (4 per lane)
__global__ void __launch_bounds__(256) red_s(float* T, int64_t rows, int64_t n, int seq) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = warp; k < n; k += nw) {
        float* p = T + row_of(k, rows, seq) * D + lane * 4;
        atomicAdd(p, 1e-9f); atomicAdd(p + 1, 1e-9f); atomicAdd(p + 2, 1e-9f); atomicAdd(p + 3, 1e-9f);
    }
}

// mode 2: plain 16-byte stores (no read)
__global__ void __launch_bounds__(256) st_v4(float* T, int64_t rows, int64_t n, int seq) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = warp; k < n; k += nw)
        reinterpret_cast<float4*>(T + row_of(k, rows, seq) * D)[lane] = make_float4(1.f, 2.f, 3.f, 4.f);
}

// mode 3: bulk copies (TMA, 512 B per row) into a shared-memory ring, one mbarrier per slot; SLOTS rows in flight per CTA
template <int SLOTS>
__global__ void __launch_bounds__(128) read_bulk(const float* __restrict__ T, int64_t rows, int64_t n, int seq, float* sink) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)SLOTS * D * 4);
    const int t = threadIdx.x;
    if (t < SLOTS) {
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar + t);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    float acc = 0.f;
    if (t < SLOTS) {                          // thread t owns slot t: issue, wait, touch, reissue
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar + t);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ring + (size_t)t * D);
        uint32_t parity = 0;
        for (int64_t k = (int64_t)blockIdx.x * SLOTS + t; k < n; k += (int64_t)gridDim.x * SLOTS) {
            const float* src = T + row_of(k, rows, seq) * D;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(D * 4) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(src), "r"(D * 4), "r"(b) : "memory");
            uint32_t done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(b), "r"(parity) : "memory");
            parity ^= 1;
            acc += ring[(size_t)t * D + (k & 127)];
        }
    }
    if (acc == 12345.678f) *sink = acc;
}

// mode 4: bulk reduce-add (512 B per row) from shared memory, G groups outstanding per thread
__global__ void __launch_bounds__(128) red_bulk(float* T, int64_t rows, int64_t n, int seq) {
    __shared__ __align__(128) float stage[D];
    for (int i = threadIdx.x; i < D; i += blockDim.x) stage[i] = 1e-9f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(stage);
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = gt; k < n; k += nt) {
        float* dst = T + row_of(k, rows, seq) * D;
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(s), "r"(D * 4) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main(int argc, char** argv) {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    printf("devices: %d\n", ndev);
    const int64_t rows = (argc > 1) ? atoll(argv[1]) : 12500000;      // 6.4 GB
    const int64_t n = (argc > 2) ? atoll(argv[2]) : (1 << 21);
    const int quick = (argc > 3) ? atoi(argv[3]) : 0;
    float *local = nullptr, *peer = nullptr, *sink = nullptr;
    CK(cudaSetDevice(0));
    CK(cudaMalloc(&local, (size_t)rows * D * 4));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(local, 0, (size_t)rows * D * 4));
    if (ndev > 1) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, 0, 1));
        printf("device 0 can access device 1: %d\n", can);
        int perf = 0, atom = 0;
        cudaDeviceGetP2PAttribute(&perf, cudaDevP2PAttrPerformanceRank, 0, 1);
        cudaDeviceGetP2PAttribute(&atom, cudaDevP2PAttrNativeAtomicSupported, 0, 1);
        printf("p2p performance rank %d, native atomics %d\n", perf, atom);
        CK(cudaDeviceEnablePeerAccess(1, 0));
        CK(cudaSetDevice(1));
        CK(cudaMalloc(&peer, (size_t)rows * D * 4));
        CK(cudaMemset(peer, 0, (size_t)rows * D * 4));
        CK(cudaDeviceSynchronize());
        CK(cudaSetDevice(0));
    }
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaFuncSetAttribute(read_bulk<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * D * 4 + 128 * 8));
    CK(cudaFuncSetAttribute(read_bulk<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * D * 4 + 32 * 8));
    struct Where { const char* name; float* T; };
    std::vector<Where> places = {{"local", local}};
    if (peer) places.push_back({"peer ", peer});
    const double gb = (double)n * D * 4 / 1e9;
    for (const Where& w : places) {
        for (int range = 0; range < 5; ++range) {
            // range 0: random over the whole table; 1: random inside the first 64 MB; 2: sequential rows;
            // 3: every access to ONE row; 4: Zipf(1) popularity over the whole table
            if (quick && (range == 1 || range == 2)) continue;
            const int64_t r = (range == 1) ? (64ll << 20) / (D * 4) : rows;
            const int seq = range == 2 ? 1 : range == 3 ? 2 : range == 4 ? 3 : 0;
            const char* rn = range == 0 ? "random rows, whole table" : range == 1 ? "random rows, 64 MB window"
                           : range == 2 ? "sequential rows" : range == 3 ? "ONE row" : "Zipf(1) rows";
            auto run = [&](const char* what, auto launch) {
                launch(); CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(a));
                for (int i = 0; i < 3; ++i) launch();
                CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
                CK(cudaGetLastError());
                const float ms = time_ms(a, b) / 3;
                printf("%s | %-26s | %-44s %8.3f ms  %7.1f GB/s  %6.1f M rows/s\n", w.name, rn, what, ms, gb / ms * 1e3, n / ms / 1e3);
                fflush(stdout);
            };
            if (!quick) {
                run("LDG.128, 2 rows in flight/warp, 32 warps/SM", [&] { read_ldg<2><<<sms * 4, 256>>>(w.T, r, n, seq, sink); });
                run("LDG.128, 8 rows in flight/warp, 64 warps/SM", [&] { read_ldg<8><<<sms * 8, 256>>>(w.T, r, n, seq, sink); });
                run("bulk copy 512 B, 32 rows in flight/CTA x4", [&] { read_bulk<32><<<sms * 4, 128, 32 * D * 4 + 32 * 8>>>(w.T, r, n, seq, sink); });
            }
            run("LDG.128, 8 rows in flight/warp, 32 warps/SM", [&] { read_ldg<8><<<sms * 4, 256>>>(w.T, r, n, seq, sink); });
            run("bulk copy 512 B, 128 rows in flight/CTA x2", [&] { read_bulk<128><<<sms * 2, 128, 128 * D * 4 + 128 * 8>>>(w.T, r, n, seq, sink); });
            run("RED.v4.f32 rows", [&] { red_v4<<<sms * 8, 256>>>(w.T, r, n, seq); });
            run("scalar RED.f32 rows", [&] { red_s<<<sms * 8, 256>>>(w.T, r, n, seq); });
            if (!quick) run("ST.128 rows", [&] { st_v4<<<sms * 8, 256>>>(w.T, r, n, seq); });
            run("bulk reduce-add 512 B", [&] { red_bulk<<<sms * 8, 128>>>(w.T, r, n, seq); });
        }
    }
    if (peer) {
        // both directions at once: device 0 works on device 1's table while device 1 works on device 0's
        CK(cudaSetDevice(1));
        CK(cudaDeviceEnablePeerAccess(0, 0));
        cudaEvent_t a1, b1;
        CK(cudaEventCreate(&a1)); CK(cudaEventCreate(&b1));
        for (int seq : {0, 3}) {
            for (int what = 0; what < 3; ++what) {
                auto go = [&](int dev, float* T) {
                    CK(cudaSetDevice(dev));
                    if (what == 0) read_ldg<8><<<sms * 4, 256>>>(T, rows, n, seq, nullptr);
                    else if (what == 1) red_v4<<<sms * 8, 256>>>(T, rows, n, seq);
                    else { read_ldg<8><<<sms * 4, 256>>>(T, rows, n, seq, nullptr); red_v4<<<sms * 8, 256>>>(T, rows, n, seq); }
                };
                go(0, peer); go(1, local);
                CK(cudaSetDevice(0)); CK(cudaDeviceSynchronize()); CK(cudaSetDevice(1)); CK(cudaDeviceSynchronize());
                CK(cudaSetDevice(0)); CK(cudaEventRecord(a)); CK(cudaSetDevice(1)); CK(cudaEventRecord(a1));
                for (int i = 0; i < 3; ++i) { go(0, peer); go(1, local); }
                CK(cudaSetDevice(0)); CK(cudaEventRecord(b)); CK(cudaSetDevice(1)); CK(cudaEventRecord(b1));
                CK(cudaSetDevice(0)); CK(cudaEventSynchronize(b)); CK(cudaSetDevice(1)); CK(cudaEventSynchronize(b1));
                float m1; CK(cudaEventElapsedTime(&m1, a1, b1));
                CK(cudaSetDevice(0));
                printf("both directions at once | %-12s | %-28s dev0 %8.3f ms  dev1 %8.3f ms per launch set\n",
                       seq ? "Zipf(1) rows" : "random rows", what == 0 ? "LDG.128 x8" : what == 1 ? "RED.v4" : "LDG.128 x8 then RED.v4",
                       time_ms(a, b) / 3, m1 / 3);
                fflush(stdout);
            }
        }
    }
    return 0;
}
