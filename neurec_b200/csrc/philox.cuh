// Philox4x32-10 (Salmon et al., SC'11) and the rejection draw built on it.
#pragma once
#include "common.cuh"

namespace nrc {

struct Philox4 {
    uint32_t x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// k-th 64-bit word of output element `elem`
__device__ __forceinline__ uint64_t philox_word(uint64_t elem, uint32_t k, uint64_t seed, uint64_t stream_id) {
    const Philox4 r = philox4x32_10((uint32_t)elem, (uint32_t)(elem >> 32), k >> 1,
                                    (uint32_t)stream_id, (uint32_t)seed,
                                    (uint32_t)(seed >> 32) ^ (uint32_t)(stream_id >> 32));
    return (k & 1u) ? (((uint64_t)r.w << 32) | r.z) : (((uint64_t)r.y << 32) | r.x);
}

// k-th candidate of output element `elem`: 64 random bits reduced modulo `high`
// (random_choice.pyx:53 `a = llrand() % c_high`).
__device__ __forceinline__ int32_t philox_candidate(uint64_t elem, uint32_t k, uint64_t seed,
                                                    uint64_t stream_id, int32_t high) {
    return (int32_t)(philox_word(elem, k, seed, stream_id) % (uint64_t)high);
}

// First candidate not contained in the sorted exclusion row; -1 when the row excludes
// everything (random_choice.pyx:32-33 raises ValueError there).
__device__ __forceinline__ int32_t philox_draw_excluding(uint64_t elem, uint64_t seed,
                                                         uint64_t stream_id, int32_t high,
                                                         const int32_t* __restrict__ excl,
                                                         int64_t deg) {
    if (deg >= high) return -1;
    for (uint32_t k = 0;; ++k) {
        const int32_t a = philox_candidate(elem, k, seed, stream_id, high);
        if (deg == 0 || !sorted_contains(excl, deg, a)) return a;
    }
}

}  // namespace nrc
