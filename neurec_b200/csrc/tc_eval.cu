// Tensor-core (tcgen05 / TMEM) candidate pass of the evaluator for large catalogues
// (BASELINE config 4: 1 M users x 10 M items x d=128, SURVEY.md 7.1 "tensor-core path").
//
// Replaces the score step of evaluator/backend/cpp/uni_evaluator.py:134 (model.predict = U.V^T,
// MF.py:120-122) for catalogues where an fp32 SIMT contraction is 30x off the machine's
// throughput.  The ranking still has to be the reference's, bit for bit, so the tensor cores
// only SELECT: scores are computed from bf16 copies of the tables (exact products, fp32
// accumulation in TMEM), every item whose approximate score can still belong to the exact top
// K+1 (a rigorous margin, see tc_prepare_users_kernel) becomes a candidate, and the candidates are
// re-scored with the oracle's fp32 FMA chain and ranked by the tie-aware selection of evaluator.cu
// (users with ties: second pass with the reference's heap root as threshold + libstdc++ heap replay).
//
// Blackwell specifics used here (sm_100a only):
//   * TMA: cp.async.bulk.tensor.2d through a CUtensorMap (SWIZZLE_128B, 128 x 64 bf16 boxes, rows
//     past the end of the table zero-filled) brings the item tiles into a ring of shared-memory
//     stages; completion is signalled with mbarrier complete_tx;
//   * tcgen05.mma.cta_group::1.kind::f16, M=128 (users) x N=128|256 (items) x K=16 per instruction,
//     issued by ONE thread; operands described by UMMA shared-memory descriptors (K-major,
//     SWIZZLE_128B: 64-element K blocks, rows 128 B apart, 16-byte chunk c of row r at c ^ (r & 7));
//   * the fp32 accumulators live in Tensor Memory (tcgen05.alloc of all 512 columns: two user
//     halves x 256 columns); tcgen05.commit -> mbarrier hands an accumulator to its epilogue warps,
//     which read it with tcgen05.ld.32x32b.x32 (warp w owns TMEM lanes 32(w%4).. = 32 users, one
//     user per thread) and release it through another mbarrier.
// The kernel, its pipelines and what bounds it are described above tc_candidate_kernel and in
// DESIGN.md section 3a; nrc_tc_gemm_debug below is the stand-alone MMA building block the tests
// use to pin the descriptor encodings (no-swizzle and SWIZZLE_128B) against a torch matmul.
#include <cuda.h>            // CUtensorMap (types only; the encoder is fetched from the driver at run time)
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace nrc {
namespace tc {

constexpr int kM = 128;        // users per CTA tile (UMMA M)
constexpr int kN = 256;        // items per tile (UMMA N)
constexpr int kUmmaK = 16;     // bf16 elements per tcgen05.mma k-step

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// TMA: one [box_rows x 64] bf16 box of a row-major [rows, K] tensor -> shared memory, laid out
// by the copy engine in the SWIZZLE_128B pattern the UMMA descriptors below expect; rows past
// the end of the tensor arrive as zeros.  Completion is signalled on `bar` (complete_tx).
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, int c_k, int c_row, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c_k), "r"(c_row), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// shared-memory writes made with ordinary stores must be fenced before the tensor core (async
// proxy) reads them
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute/arch/mma_sm100_desc.hpp bit layout):
// [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout type = 0.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type = 0) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout_type & 7u) << 61;
    return d;
}

// SWIZZLE_128B K-major layout (layout type 2): the tile is cut along K into blocks of 64 bf16
// (128 B); inside a block row r occupies 128 contiguous bytes at r*128 and its 16-byte chunk c
// sits at position c ^ (r & 7); 8-row groups are 1024 B apart (SBO), LBO is unused (1).  A warp
// copying 4 rows x 8 chunks reads 4 x 128 contiguous global bytes and stores conflict-free.
__device__ __forceinline__ void load_tile_sw128(uint8_t* smem, const __nv_bfloat16* __restrict__ g, int rows,
                                                int valid_rows, int K, int tid, int nthreads) {
    const int chunks = K >> 3;
    for (int idx = tid; idx < rows * chunks; idx += nthreads) {
        const int row = idx / chunks, cg = idx - row * chunks;
        const int blk = cg >> 3, c = cg & 7;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (row < valid_rows) v = __ldg(reinterpret_cast<const uint4*>(g + (size_t)row * K) + cg);
        *reinterpret_cast<uint4*>(smem + (size_t)blk * rows * 128 + (size_t)row * 128 + ((c ^ (row & 7)) << 4)) = v;
    }
}

// descriptors of k-step s (16 bf16) for a SWIZZLE_128B tile of `rows` rows
__device__ __forceinline__ uint64_t sw128_desc(uint32_t tile_base, int rows, int s) {
    return make_smem_desc(tile_base + (uint32_t)(s >> 2) * rows * 128 + (uint32_t)(s & 3) * 32, 16, 1024, 2);
}

// Instruction descriptor for kind::f16: D = f32 (c_format 1 @ bit 4), A = B = bf16 (format 1 @ bits
// 7 and 10), both K-major (bits 15, 16 = 0), N >> 3 @ bit 17, M >> 4 @ bit 24.
__host__ __device__ constexpr uint32_t make_instr_desc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread (lane) gets its row's columns [c, c+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// The same load without the wait, so that the next chunk can be in flight while this one is used.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// tcgen05.wait::ld; the registers are in/out operands so that no use of them is scheduled above it
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                   "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                   "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// Copy a [rows, K] row-major bf16 tile from global memory into the canonical no-swizzle K-major
// layout: 16-byte chunk (row, kc) -> smem[(kc * rows + row) * 16 B].  Rows >= valid_rows are zero.
__device__ __forceinline__ void load_tile_kmajor(uint8_t* smem, const __nv_bfloat16* __restrict__ g, int rows,
                                                 int valid_rows, int K, int tid, int nthreads) {
    const int chunks = K >> 3;   // 8 bf16 per 16-byte chunk
    for (int idx = tid; idx < rows * chunks; idx += nthreads) {
        const int row = idx % rows, kc = idx / rows;   // consecutive threads -> consecutive rows: conflict-free smem
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (row < valid_rows) v = __ldg(reinterpret_cast<const uint4*>(g + (size_t)row * K) + kc);
        *reinterpret_cast<uint4*>(smem + ((size_t)kc * rows + row) * 16) = v;
    }
}

// ----------------------------------------------------------------------------------------
// Stage-1 self-test kernel: out[128, 256] = A[128, K] . B[256, K]^T through tcgen05/TMEM.
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(192)
tc_gemm_debug_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, int K,
                     float* __restrict__ out, int swizzle) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                          // [K/8][128] x 16 B
    uint8_t* sB = sA + (size_t)kM * K * 2;       // [K/8][256] x 16 B
    __shared__ uint64_t bar_full;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (swizzle) {
        load_tile_sw128(sA, A, kM, kM, K, tid, blockDim.x);
        load_tile_sw128(sB, B, kN, kN, K, tid, blockDim.x);
    } else {
        load_tile_kmajor(sA, A, kM, kM, K, tid, blockDim.x);
        load_tile_kmajor(sB, B, kN, kN, K, tid, blockDim.x);
    }
    fence_async_smem();
    if (tid == 0) {
        mbar_init(&bar_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {   // one warp allocates (and later frees) the accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"((uint32_t)kN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_slot;

    if (warp == 4 && lane == 0) {
        const uint32_t idesc = make_instr_desc(kM, kN);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        for (int s = 0; s < K / kUmmaK; ++s) {
            const uint64_t ad = swizzle ? sw128_desc(a0, kM, s) : make_smem_desc(a0 + s * 2 * kM * 16, kM * 16, 128);
            const uint64_t bd = swizzle ? sw128_desc(b0, kN, s) : make_smem_desc(b0 + s * 2 * kN * 16, kN * 16, 128);
            umma_bf16(tmem_d, ad, bd, idesc, s > 0 ? 1u : 0u);
        }
        umma_commit(&bar_full);
    }
    if (warp < 4) {
        mbar_wait(&bar_full, 0);
        tc_fence_after();
        const int row = warp * 32 + lane;
        for (int c = 0; c < kN; c += 32) {
            float v[32];
            tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) out[(size_t)row * kN + c + i] = v[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)kN));
}

}  // namespace tc
}  // namespace nrc

using namespace nrc;

// Test hook: out f32 [128, 256] = A bf16 [128, K] . B bf16 [256, K]^T  (K multiple of 16, <= 256).
extern "C" int nrc_tc_gemm_debug(const void* a_bf16, const void* b_bf16, int32_t k, int32_t swizzle, float* out,
                                 void* stream) {
    NRC_REQUIRE(k >= 16 && k <= 256 && (k % 16) == 0, NRC_E_LIMIT, "k must be a multiple of 16 in [16, 256]");
    NRC_REQUIRE(!swizzle || (k % 64) == 0, NRC_E_LIMIT, "the SWIZZLE_128B layout needs k % 64 == 0");
    const size_t smem = (size_t)(tc::kM + tc::kN) * k * 2;
    NRC_CUDA_CHECK(cudaFuncSetAttribute(tc::tc_gemm_debug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    tc::tc_gemm_debug_kernel<<<1, 192, smem, as_stream(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(a_bf16), reinterpret_cast<const __nv_bfloat16*>(b_bf16), k, out,
        swizzle);
    NRC_CUDA_CHECK(cudaGetLastError());
    return NRC_OK;
}

// =========================================================================================
// Candidate pass + finalisation
// =========================================================================================
namespace nrc {
namespace tc {

// Candidate kernel.  One CTA owns 256 users (two UMMA M=128 halves, both multiplied with every
// item tile, so the bf16 item table is read from L2 once per 256 users -- with 128 users per CTA
// the kernel sat on the L2->SM bandwidth cap) and one contiguous range of item tiles.
//   warps 0-7  epilogue: warp w serves user half h = w / 4 through TMEM lanes 32*(w%4) .. +31;
//              one thread = one user: running threshold, min-heap of the LQ best approximate
//              scores, merge-walk over the user's train row, candidate list
//   warp 8     MMA issuer (one thread): 2 x D/16 tcgen05.mma (M128 N{256|128} K16) per item tile
//   warp 9     TMA producer (one thread): item tile -> shared memory (SWIZZLE_128B boxes)
// Pipelines: `nst` shared-memory item stages (full_b / empty_b) and, per user half, 256/NT TMEM
// accumulators of NT fp32 columns (acc_full / acc_empty per half): the MMAs of one half overlap
// the filtering of the other, and with NT = 128 also the filtering of the same half's previous tile.
// An SS-mode M128 N128 K16 MMA reads 8 KB of shared memory per 64 clk, exactly the 128 B/clk
// shared-memory limit (N256 needs 96 B/clk), but the kernel as a whole is paced by the epilogue's
// Tensor Memory read-out, so both tile widths measure the same; see DESIGN.md 3a.
constexpr int kMaxList = 64;           // threshold rank <= 64 (2*top_k for the tie-replay pass)
constexpr int kMU = 256;               // users per CTA
constexpr int kMaxStages = 3;
constexpr int kCandThreads = 320;          // CH = 1: 8 epilogue warps + MMA + TMA; CH = 2: 16 epilogue warps (576 threads)

struct CandArgs {
    const __nv_bfloat16* Ub;   // [num_eval, D] bf16 rows of the users being evaluated (gathered)
    const __nv_bfloat16* Vb;   // [N, D] bf16 item table
    const float* margin;       // [num_eval] 2 * eps_u (see tc_prepare_users_kernel)
    const int32_t* users;      // [num_eval] user ids (train CSR is indexed by user id)
    const int64_t* train_ptr; const int32_t* train_idx;
    int dbg;                   // NRC_TC_DBG experiment bits (0 in normal use): 1 skip epilogue, 2 skip TMA, 4 skip MMA, 8 TMEM read-out only
    int num_eval, N, D;
    int LQ;                    // rank of the running threshold kept per (user, list)
    int lstride;               // shared-memory words per list (odd: conflict-free whatever entry a lane touches)
    int nst;                   // shared-memory item stages (2 or 3)
    int seg_tiles;             // item tiles per grid.y segment
    int nslots;                // candidate lists per user = gridDim.y * CH
    int cap;                   // entries per list
    int32_t* cand;             // [num_eval, nslots, cap] candidate item ids, ascending inside a list
    float* cand_val;           // same shape: the approximate (bf16 tensor-core) score of each candidate
    int32_t* cand_cnt;         // [num_eval, nslots] candidates seen (> cap => overflow)
};

// CH = column halves per tile with their own epilogue warps: 1 -> 8 epilogue warps, one user per
// thread over all NT columns; 2 -> 16 epilogue warps, a user is served by two threads (columns
// [0, NT/2) and [NT/2, NT)), each with its OWN threshold list and candidate list (slot), like two
// item segments interleaved -- twice the warps to hide the tcgen05.ld -> filter dependency chains.
template <int NT, int CH>   // items per tile (UMMA N): 128 (default) or 256
__global__ void __launch_bounds__((8 * CH + 2) * 32, 1)
tc_candidate_kernel(const CandArgs P, const __grid_constant__ CUtensorMap tmapV) {
    constexpr int kMmaWarp = 8 * CH, kTmaWarp = 8 * CH + 1;
    constexpr int kThreads = (8 * CH + 2) * 32;
    constexpr int kAccStages = 256 / NT;   // TMEM stages per user half: 512 columns = 2 halves x kAccStages x NT
    extern __shared__ __align__(1024) uint8_t smem[];
    const int D = P.D;
    const uint32_t half_bytes = (uint32_t)kM * D * 2;     // one 128-user operand tile
    const uint32_t stage_bytes = (uint32_t)NT * D * 2;   // one 128-item operand tile
    uint8_t* sA = smem;                                   // 2 x [128 x D] bf16, SWIZZLE_128B blocks
    uint8_t* sB = sA + 2 * (size_t)half_bytes;            // nst x [128 x D] bf16
    float* sList = reinterpret_cast<float*>(sB + (size_t)P.nst * stage_bytes);   // [256][lstride]
    __shared__ uint64_t full_b[kMaxStages], empty_b[kMaxStages], acc_full[4], acc_empty[4];   // [stage * 2 + half]
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * kMU;
    const int t_begin = blockIdx.y * P.seg_tiles;                         // first item tile of this CTA
    const int T = min(P.seg_tiles, (P.N + NT - 1) / NT - t_begin);      // its tile count (>= 1)

    for (int h = 0; h < 2; ++h)
        load_tile_sw128(sA + (size_t)h * half_bytes, P.Ub + (size_t)(row0 + h * kM) * D, kM,
                        max(0, min(kM, P.num_eval - row0 - h * kM)), D, tid, kThreads);
    for (int i = tid; i < CH * kMU * P.lstride; i += kThreads) sList[i] = -INFINITY;
    fence_async_smem();
    if (tid == 0) {
        for (int i = 0; i < kMaxStages; ++i) {
            mbar_init(&full_b[i], 1);    // the producer's arrive.expect_tx; the copy engine completes the bytes
            mbar_init(&empty_b[i], 1);   // tcgen05.commit of the tile that read the stage
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 128 * CH);   // the four (eight) epilogue warps of one user half
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == kTmaWarp) {
        // ---------------- producer (one thread): item tiles -> shared memory by TMA ----------------
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapV)) : "memory");
            const uint32_t b0 = smem_u32(sB);
            int s = 0, ph = 0;
            for (int t = 0; t < T; ++t) {
                mbar_wait(&empty_b[s], ph ^ 1);
                if (P.dbg & 2) {
                    mbar_arrive(&full_b[s]);
                } else {
                    mbar_arrive_expect_tx(&full_b[s], stage_bytes);
                    for (int kb = 0; kb < D / 64; ++kb)   // one 128 x 64 box per 128-byte K block
                        tma_load_2d(b0 + (uint32_t)s * stage_bytes + (uint32_t)kb * NT * 128, &tmapV, kb * 64,
                                    (t_begin + t) * NT, &full_b[s]);
                }
                if (++s == P.nst) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == kMmaWarp) {
        // ---------------- MMA issuer (one thread) ----------------
        if (lane == 0) {
            const uint32_t idesc = make_instr_desc(kM, NT);
            // Descriptors differ only in the 14-bit start-address field (16-byte units; shared
            // memory is < 256 KB, so adding offsets never carries out of the field).
            const uint64_t a_base = sw128_desc(smem_u32(sA), kM, 0);
            const uint64_t b_base = sw128_desc(smem_u32(sB), NT, 0);
            const int nks = D / kUmmaK;
            int s = 0, ph = 0;
            for (int t = 0; t < T; ++t) {
                const int a = t % kAccStages, pa = (t / kAccStages) & 1;
                mbar_wait(&full_b[s], ph);
                const uint64_t bb = b_base + (uint64_t)(((uint32_t)s * stage_bytes) >> 4);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    mbar_wait(&acc_empty[a * 2 + h], pa ^ 1);
                    tc_fence_after();
                    if (!(P.dbg & 4)) {
                        const uint32_t td = tmem_base + (uint32_t)((a * 2 + h) * NT);
                        const uint64_t ah = a_base + (uint64_t)(((uint32_t)h * half_bytes) >> 4);
#pragma unroll 4
                        for (int ks = 0; ks < nks; ++ks) {
                            // k-step ks: 64-element K block (ks >> 2), 32 B per step inside it; the block
                            // pitch is rows x 128 B, which differs between the user and the item tile
                            const uint32_t kblk = (uint32_t)ks >> 2, kin = ((uint32_t)ks & 3u) * 32u;
                            umma_bf16(td, ah + (uint64_t)((kblk * kM * 128u + kin) >> 4),
                                      bb + (uint64_t)((kblk * NT * 128u + kin) >> 4), idesc, ks > 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(&acc_full[a * 2 + h]);   // this half's 128 x NT accumulator is complete
                }
                umma_commit(&empty_b[s]);    // the shared-memory stage may be refilled
                if (++s == P.nst) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // ---------------- epilogue: one user per thread ----------------
        const int h = (warp >> 2) & 1, wq = warp & 3;      // user half and TMEM lane quarter
        const int ch = warp >> 3;                          // column half served by this warp (0 when CH == 1)
        constexpr int CW = NT / CH;                        // columns per epilogue thread and tile
        const int r = h * kM + wq * 32 + lane;       // user row inside the CTA
        const int row = row0 + r;
        const bool live = row < P.num_eval;
        float* lst = sList + (ch * kMU + r) * P.lstride;
        const int slot = blockIdx.y * CH + ch;
        const float margin = live ? P.margin[row] : 0.0f;
        const int u = live ? P.users[row] : 0;
        const int64_t tb = live ? P.train_ptr[u] : 0;
        const int tl = live ? (int)(P.train_ptr[u + 1] - tb) : 0;
        int32_t* my_cand = P.cand + ((size_t)row * P.nslots + slot) * P.cap;
        float* my_val = P.cand_val + ((size_t)row * P.nslots + slot) * P.cap;
        float thr = -INFINITY, thr_m = live ? -INFINITY : INFINITY;   // padding rows never produce candidates
        int cnt = 0;
        // merge-walk over the user's sorted train row: items arrive in ascending order, so the
        // mask test of a candidate is "advance the cursor to >= item, compare" (amortised O(deg))
        int tpos = 0;
        if (t_begin > 0) {   // first train item at or after this segment's first item
            int lo = 0, hi = tl;
            const int first = t_begin * NT;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(P.train_idx + tb + mid) < first) lo = mid + 1; else hi = mid;
            }
            tpos = lo;
        }
        int tnext = (tpos < tl) ? __ldg(P.train_idx + tb + tpos) : INT32_MAX;
        int tahead = (tpos + 1 < tl) ? __ldg(P.train_idx + tb + tpos + 1) : INT32_MAX;   // prefetched: advancing never waits on memory
        for (int t = 0; t < T; ++t) {
            const int a = t % kAccStages, pa = (t / kAccStages) & 1;
            const bool tail = (t_begin + t + 1) * NT > P.N;   // only the catalogue's last tile has columns past N
            mbar_wait(&acc_full[a * 2 + h], pa);
            tc_fence_after();
            // filter one chunk of 32 columns (items (t_begin + t)*NT + c ..) held in registers
            auto filter_chunk = [&](const uint32_t (&raw)[32], const int c) {
                if (P.dbg & 8) return;   // experiment: Tensor Memory read-out only
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
                // cheap common case: the chunk maximum does not reach the threshold
                float t16[16];   // max tree: few warps per scheduler, so dependent chains are exposed
#pragma unroll
                for (int i = 0; i < 16; ++i) t16[i] = fmaxf(v[2 * i], v[2 * i + 1]);
#pragma unroll
                for (int i = 0; i < 8; ++i) t16[i] = fmaxf(t16[2 * i], t16[2 * i + 1]);
#pragma unroll
                for (int i = 0; i < 4; ++i) t16[i] = fmaxf(t16[2 * i], t16[2 * i + 1]);
                const float mx = fmaxf(fmaxf(t16[0], t16[1]), fmaxf(t16[2], t16[3]));
                unsigned m = 0u;
                if (mx > thr_m) {   // bit i set when column i can still matter for this user
                    unsigned m4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int i = 0; i < 32; ++i) m4[i & 3] |= (v[i] > thr_m ? 1u : 0u) << i;
                    m = (m4[0] | m4[1]) | (m4[2] | m4[3]);
                }
                const int item0 = (t_begin + t) * NT + c;
                if (tail && item0 + 32 > P.N) m &= (item0 >= P.N) ? 0u : ((1u << (P.N - item0)) - 1u);
                while (m) {                       // rare: ~LQ ln(N/LQ) times per user in total
                    const int i = __ffs(m) - 1;
                    m &= m - 1;
                    float x = v[0];
#pragma unroll
                    for (int j = 1; j < 32; ++j) x = (j == i) ? v[j] : x;   // register select, no local memory
                    if (!(x > thr_m)) continue;   // the threshold may have risen inside this chunk
                    const int item = item0 + i;
                    while (tnext < item) {
                        ++tpos;
                        tnext = tahead;
                        tahead = (tpos + 1 < tl) ? __ldg(P.train_idx + tb + tpos + 1) : INT32_MAX;
                    }
                    if (tnext == item) continue;  // train item: masked to -inf by the reference
                    if (cnt < P.cap) { my_cand[cnt] = item; my_val[cnt] = x; }
                    ++cnt;
                    if (x > thr) {                // keep the LQ best approximate scores in a min-heap
                        int hpos = 0;             // replace the root (the LQ-th best) and sift down
                        for (;;) {
                            int ch = 2 * hpos + 1;
                            if (ch >= P.LQ) break;
                            float cv = lst[ch];
                            if (ch + 1 < P.LQ) {
                                const float cv2 = lst[ch + 1];
                                if (cv2 < cv) { cv = cv2; ++ch; }
                            }
                            if (!(cv < x)) break;
                            lst[hpos] = cv;
                            hpos = ch;
                        }
                        lst[hpos] = x;
                        thr = lst[0];
                        thr_m = thr - margin;
                    }
                }
            };
            if (P.dbg & 1) { tc_fence_before(); mbar_arrive(&acc_empty[a * 2 + h]); continue; }
            // two chunks in flight: tcgen05.ld of chunk c+1 overlaps the filtering of chunk c
            const uint32_t tbase = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)((a * 2 + h) * NT);
            uint32_t ra[32], rb[32];
            __syncwarp();   // the candidate branch diverges; tcgen05.ld needs the whole warp
            const int c_beg = ch * CW, c_end = c_beg + CW;
            tmem_ld32_issue(tbase + (uint32_t)c_beg, ra);
#pragma unroll 1
            for (int c = c_beg; c < c_end; c += 64) {
                tmem_ld_wait(ra);
                tmem_ld32_issue(tbase + (uint32_t)(c + 32), rb);
                filter_chunk(ra, c);
                __syncwarp();
                tmem_ld_wait(rb);
                if (c + 64 < c_end) tmem_ld32_issue(tbase + (uint32_t)(c + 64), ra);
                filter_chunk(rb, c + 32);
                __syncwarp();
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[a * 2 + h]);
        }
        if (live) P.cand_cnt[(size_t)row * P.nslots + slot] = cnt;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

// bf16 copy of the item table + largest row norm (positive floats order like their bit patterns)
__global__ void tc_prepare_items_kernel(const float* __restrict__ V, int64_t N, int D, __nv_bfloat16* __restrict__ Vb,
                                        unsigned int* __restrict__ vmax_bits) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float best = 0.0f;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < N; i += warps) {
        float sq = 0.0f;
        for (int k = lane; k < D; k += 32) {
            const float x = V[i * D + k];
            Vb[i * D + k] = __float2bfloat16_rn(x);
            sq = fmaf(x, x, sq);
        }
        sq = warp_sum(sq);
        best = fmaxf(best, sqrtf(sq));
    }
    if (lane == 0) atomicMax(vmax_bits, __float_as_uint(best));
}

// bf16 rows of the evaluated users + their candidate margin
//   bf16 keeps 8 significant bits, so round-to-nearest has unit roundoff 2^-8 per FACTOR:
//   u^_k v^_k = u_k v_k (1+a)(1+b), |a|,|b| <= 2^-8  =>  |u^.v^ - u.v| <= (2^-7 + 2^-16) sum|u_k v_k|
//   <= (2^-7 + 2^-16) |u| |v| (Cauchy-Schwarz).  The fp32 accumulation in Tensor Memory (truncating
//   adder, d <= 192 terms) and the fp32 FMA chain of the exact score add < 2^-11 |u| |v| together.
//   eps = (2^-7 + 2^-11) * |u| * max_i |v_i| * 1.001 (norms are fp32-rounded), margin = 2 * eps:
//   an item's approximate score and the order statistic it is compared with each move by <= eps.
__global__ void tc_prepare_users_kernel(const float* __restrict__ U, const int32_t* __restrict__ users, int num_eval,
                                        int D, const unsigned int* __restrict__ vmax_bits,
                                        __nv_bfloat16* __restrict__ Ub, float* __restrict__ margin) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= num_eval) return;
    const float* u = U + (size_t)users[row] * D;
    float sq = 0.0f;
    for (int k = lane; k < D; k += 32) {
        const float x = u[k];
        Ub[(size_t)row * D + k] = __float2bfloat16_rn(x);
        sq = fmaf(x, x, sq);
    }
    sq = warp_sum(sq);
    if (lane == 0) {
        const float vmax = __uint_as_float(*vmax_bits);
        margin[row] = 2.0f * (0.0078125f + 0.00048828125f) * sqrtf(sq) * vmax * 1.001f;
    }
}

}  // namespace tc
}  // namespace nrc

#include "tc_eval.cuh"

namespace nrc {
namespace tc {

// library-owned workspaces, grown on demand (never inside a stream capture)
struct Arena {
    void* p = nullptr;
    size_t bytes = 0;
    int reserve(size_t need) {
        if (need <= bytes) return NRC_OK;
        if (p) NRC_CUDA_CHECK(cudaFree(p));
        p = nullptr; bytes = 0;
        NRC_CUDA_CHECK(cudaMalloc(&p, need));
        bytes = need;
        return NRC_OK;
    }
};
static Arena g_items;          // bf16 item table + max row norm
static Arena g_pass[2];        // per pass: bf16 user rows, margins, candidate lists and counts
static const __nv_bfloat16* g_vb = nullptr;
static unsigned int* g_vmax = nullptr;
static int g_items_n = 0, g_items_d = 0;
static cudaEvent_t g_ev[2] = {nullptr, nullptr};   // around the last main-pass tc_candidate_kernel launch
static double g_last_flops = 0.0;

static size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

// The bf16 copy + max row norm are reused across calls while the caller vouches that the table has
// not changed: nrc_eval_tc_items_version(v != 0) keys the cache on (pointer, shape, v); version 0
// (default) converts on every call.
static int g_ch_pref = -1;      // epilogue layout: 1 = 8 warps, 2 = 16 warps (nrc_eval_tc_epilogue_warps / NRC_TC_CH)
static uint64_t g_items_version = 0, g_cached_version = 0;
static const float* g_cached_ptr = nullptr;

int prepare_items(const float* V, int D, int N, cudaStream_t st) {
    NRC_REQUIRE(D % 64 == 0 && D >= 64 && D <= 256, NRC_E_LIMIT,
                "the tensor-core pass needs dim in {64, 128, 192, 256} (got %d)", D);
    if (g_items_version != 0 && g_cached_version == g_items_version && g_cached_ptr == V && g_items_n == N &&
        g_items_d == D && g_vb != nullptr)
        return NRC_OK;
    const size_t o_vmax = up256((size_t)N * D * 2);
    int rc = g_items.reserve(o_vmax + 256);
    if (rc) return rc;
    uint8_t* ws = reinterpret_cast<uint8_t*>(g_items.p);
    __nv_bfloat16* Vb = reinterpret_cast<__nv_bfloat16*>(ws);
    g_vmax = reinterpret_cast<unsigned int*>(ws + o_vmax);
    NRC_CUDA_CHECK(cudaMemsetAsync(g_vmax, 0, 4, st));
    tc_prepare_items_kernel<<<sm_count() * 8, 256, 0, st>>>(V, N, D, Vb, g_vmax);
    NRC_CUDA_CHECK(cudaGetLastError());
    g_vb = Vb; g_items_n = N; g_items_d = D;
    g_cached_ptr = V; g_cached_version = g_items_version;
    return NRC_OK;
}

int run_pass(int pass, const float* U, const int32_t* users, int num_rows, const int64_t* train_ptr,
             const int32_t* train_idx, int LQ, int cap, CandLists* out, cudaStream_t st) {
    NRC_REQUIRE(pass == 0 || pass == 1, NRC_E_VALUE, "pass must be 0 or 1");
    NRC_REQUIRE(g_vb != nullptr, NRC_E_VALUE, "prepare_items has not run");
    NRC_REQUIRE(LQ >= 1 && LQ <= kMaxList, NRC_E_LIMIT, "threshold rank %d outside [1, %d]", LQ, kMaxList);
    const int D = g_items_d, N = g_items_n;
    const int row_tiles = (num_rows + kMU - 1) / kMU;
    // Tile width.  Both instantiations end up paced by the Tensor Memory read-out of the epilogue
    // (every fp32 accumulator is read once: 128 KB per 128 x 256 block at ~64 B/clk/SM), so the
    // default is the 128-item tile, whose double-buffered accumulators tolerate bursts of
    // candidates better; NRC_TC_NT=256 selects the 256-item tile (fewer shared-memory reads per
    // MMA, single-buffered accumulators) when two such stages fit.
    const int lstride = (LQ <= 32) ? 33 : 65;
    // 16 epilogue warps (two column halves per user, NRC_TC_CH=2) when their lists fit beside >= 2 stages
    if (g_ch_pref < 0) { const char* e = getenv("NRC_TC_CH"); g_ch_pref = e ? atoi(e) : 1; }
    const int CH = (g_ch_pref == 2 && pass == 0 &&
              (size_t)2 * kM * D * 2 + (size_t)2 * kMU * lstride * 4 + 2048 + (size_t)2 * 128 * D * 2 <= 227 * 1024) ? 2 : 1;
    const size_t list_bytes = (size_t)kMU * CH * lstride * 4;
    const size_t fixed = (size_t)2 * kM * D * 2 + list_bytes + 2048;
    const size_t budget = 227 * 1024 - fixed;
    const char* nt_env = getenv("NRC_TC_NT");
    int kNT = 128;
    if (nt_env && atoi(nt_env) == 256 && budget / ((size_t)256 * D * 2) >= 2) kNT = 256;
    const int T = (N + kNT - 1) / kNT;
    // Item segments (grid.y): with few user tiles, split the catalogue so that every SM has a CTA.
    // Each segment restarts its threshold (still a lower bound of the true one), which costs a few
    // more candidates; pick the split with the fewest waves per unit of work, at most 8 unless a
    // single wave needs more (16 at most; 48 in the replay pass), and never segments shorter
    // than 4096 items.
    int G = 1;
    {
        const int sms = sm_count();
        const int few = (pass == 1) ? 48 : 16;   // the replay pass re-scores its lists with a whole CTA per user
        const int gmax = (row_tiles * 8 < sms) ? ((sms / row_tiles < few) ? sms / row_tiles : few) : 8;
        double best = 1e30;
        const int min_tiles = 4096 / kNT;
        for (int g = 1; g <= gmax && g * min_tiles <= (T > min_tiles ? T : min_tiles); ++g) {
            const int ctas = row_tiles * g;
            const double cost = (double)((ctas + sms - 1) / sms) / g * (1.0 + 0.02 * (g - 1));
            if (cost < best - 1e-9) { best = cost; G = g; }
        }
    }
    const int seg_tiles = (T + G - 1) / G;
    G = (T + seg_tiles - 1) / seg_tiles;           // no empty segment
    const int nslots = G * CH;
    int nst = (int)(budget / ((size_t)kNT * D * 2));
    if (nst > kMaxStages) nst = kMaxStages;
    NRC_REQUIRE(nst >= 2, NRC_E_LIMIT, "dim %d with threshold rank %d does not fit the tensor-core pass", D, LQ);
    const size_t rows_pad = (size_t)row_tiles * kMU;
    const size_t o_ub = 0;
    const size_t o_margin = o_ub + up256(rows_pad * D * 2);
    const size_t o_cnt = o_margin + up256(rows_pad * 4);
    const size_t o_cand = o_cnt + up256(rows_pad * nslots * 4);
    const size_t o_scr = o_cand + up256(rows_pad * (size_t)nslots * cap * 4);
    const size_t total = o_scr + up256(rows_pad * (size_t)nslots * cap * 4);
    int rc = g_pass[pass].reserve(total);
    if (rc) return rc;
    uint8_t* ws = reinterpret_cast<uint8_t*>(g_pass[pass].p);
    __nv_bfloat16* Ub = reinterpret_cast<__nv_bfloat16*>(ws + o_ub);
    float* margin = reinterpret_cast<float*>(ws + o_margin);
    int32_t* cnt = reinterpret_cast<int32_t*>(ws + o_cnt);
    int32_t* cd = reinterpret_cast<int32_t*>(ws + o_cand);
    tc_prepare_users_kernel<<<(num_rows * 32 + 255) / 256, 256, 0, st>>>(U, users, num_rows, D, g_vmax, Ub, margin);
    NRC_CUDA_CHECK(cudaGetLastError());

    const char* dbg_env = getenv("NRC_TC_DBG");
    CandArgs P{Ub, g_vb, margin, users, train_ptr, train_idx, dbg_env ? atoi(dbg_env) : 0, num_rows, N, D,
               LQ, lstride, nst, seg_tiles, nslots, cap, cd, reinterpret_cast<float*>(ws + o_scr), cnt};
    CUtensorMap tmapV;
    {   // bf16 item table [N, D] row-major; box = 128 items x 64 k (one 128-byte swizzle span)
        static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
        if (!encode) {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            NRC_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
            NRC_REQUIRE(fn != nullptr && q == cudaDriverEntryPointSuccess, NRC_E_CUDA,
                        "the driver does not export cuTensorMapEncodeTiled");
            encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
        }
        const cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)N};
        const cuuint64_t gstride[1] = {(cuuint64_t)D * 2};
        const cuuint32_t box[2] = {64u, (cuuint32_t)kNT};
        const cuuint32_t estr[2] = {1u, 1u};
        const CUresult cr = encode(&tmapV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)g_vb, gdim, gstride, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        NRC_REQUIRE(cr == CUDA_SUCCESS, NRC_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)cr);
    }
    size_t smem = (size_t)2 * kM * D * 2 + (size_t)nst * kNT * D * 2 + list_bytes;
    if (smem < 120 * 1024) smem = 120 * 1024;   // one CTA per SM: each CTA allocates all 512 TMEM columns
    static bool attr_done = false;
    if (!attr_done) {
        NRC_CUDA_CHECK(cudaFuncSetAttribute(tc_candidate_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            226 * 1024));
        NRC_CUDA_CHECK(cudaFuncSetAttribute(tc_candidate_kernel<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            226 * 1024));
        NRC_CUDA_CHECK(cudaFuncSetAttribute(tc_candidate_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            226 * 1024));
        NRC_CUDA_CHECK(cudaFuncSetAttribute(tc_candidate_kernel<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            226 * 1024));
        attr_done = true;
    }
    const dim3 grid(row_tiles, G);
    if (pass == 0) {
        if (!g_ev[0]) {
            NRC_CUDA_CHECK(cudaEventCreate(&g_ev[0]));
            NRC_CUDA_CHECK(cudaEventCreate(&g_ev[1]));
        }
        NRC_CUDA_CHECK(cudaEventRecord(g_ev[0], st));
    }
    if (CH == 2) {
        if (kNT == 256) tc_candidate_kernel<256, 2><<<grid, 576, smem, st>>>(P, tmapV);
        else tc_candidate_kernel<128, 2><<<grid, 576, smem, st>>>(P, tmapV);
    } else {
        if (kNT == 256) tc_candidate_kernel<256, 1><<<grid, kCandThreads, smem, st>>>(P, tmapV);
        else tc_candidate_kernel<128, 1><<<grid, kCandThreads, smem, st>>>(P, tmapV);
    }
    NRC_CUDA_CHECK(cudaGetLastError());
    if (pass == 0) {
        NRC_CUDA_CHECK(cudaEventRecord(g_ev[1], st));
        g_last_flops = 2.0 * (double)num_rows * (double)N * (double)D;
    }
    out->cand = cd;
    out->scratch = reinterpret_cast<float*>(ws + o_scr);
    out->margin = margin;
    out->cnt = cnt;
    out->nslots = nslots;
    out->cap = cap;
    return NRC_OK;
}

}  // namespace tc
}  // namespace nrc

// Evaluations of one fixed model in several calls (user batches) share the bf16 item table: set a
// non-zero version before the first call and keep it while the table is unchanged; any other value
// (or 0 = never cache) makes the next call convert again.
extern "C" int nrc_eval_tc_items_version(uint64_t version) {
    nrc::tc::g_items_version = version;
    return NRC_OK;
}

// 8 (default) or 16 epilogue warps in the candidate kernel of the main pass (16: every user is served by
// two threads, one per half of each item tile, each with its own threshold and candidate list).
extern "C" int nrc_eval_tc_epilogue_warps(int32_t warps) {
    NRC_REQUIRE(warps == 8 || warps == 16, NRC_E_VALUE, "epilogue warps must be 8 or 16 (got %d)", warps);
    nrc::tc::g_ch_pref = warps / 8;
    return NRC_OK;
}

// Duration (CUDA events on the launching stream) and algorithmic flops (2 * users * items * dim)
// of the last tcgen05 candidate-kernel launch; waits for that launch to finish.
extern "C" int nrc_eval_tc_last_launch(float* kernel_ms, double* flops) {
    NRC_REQUIRE(kernel_ms != nullptr && flops != nullptr, NRC_E_VALUE, "NULL output");
    NRC_REQUIRE(nrc::tc::g_ev[0] != nullptr, NRC_E_VALUE, "nrc_eval_mf_tc has not run yet");
    NRC_CUDA_CHECK(cudaEventSynchronize(nrc::tc::g_ev[1]));
    NRC_CUDA_CHECK(cudaEventElapsedTime(kernel_ms, nrc::tc::g_ev[0], nrc::tc::g_ev[1]));
    *flops = nrc::tc::g_last_flops;
    return NRC_OK;
}
