// Philox4x32-10 counter-based sampler for BPR negatives (K1).
//
// Replaces `choice(itemList)` + redraw-while-played of recommender/cf/BPR.py:46-49.  The draw
// for attempt t of global event e in epoch E is a pure function of (seed, E, e, slot, t), so
// the result does not depend on how events are scheduled or sharded (see oracle/philox.py for
// the definition; KATs in tests/test_oracle_golden.py).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace yue {

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                       uint32_t c3, uint32_t k0, uint32_t k1,
                                                       uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// is `x` in the sorted row [row, row+len)?
__device__ __forceinline__ bool row_contains(const int32_t* __restrict__ row, int len, int32_t x) {
    int lo = 0, hi = len;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int32_t v = __ldg(row + mid);
        if (v < x) lo = mid + 1; else hi = mid;
    }
    return lo < len && __ldg(row + lo) == x;
}

constexpr uint32_t kMaxAttempts = 1u << 22;

// One accepted negative for global event `e`.  Returns -1 only if the row covers the catalog.
__device__ __forceinline__ int32_t sample_negative(uint64_t seed, uint32_t epoch, uint64_t e,
                                                   uint32_t slot, uint32_t n_items,
                                                   const int32_t* __restrict__ row, int len) {
    uint32_t w[4];
    for (uint32_t t = 0; t < kMaxAttempts; ++t) {
        if ((t & 3u) == 0u)
            philox4x32_10((uint32_t)e, (uint32_t)(e >> 32), epoch, (slot << 20) | (t >> 2),
                          (uint32_t)seed, (uint32_t)(seed >> 32), w);
        const uint32_t s = t & 3u;   // selects, not w[s]: keeps the words in registers
        const uint32_t r = s == 0u ? w[0] : (s == 1u ? w[1] : (s == 2u ? w[2] : w[3]));
        const int32_t j = (int32_t)__umulhi(r, n_items);
        if (!row_contains(row, len, j)) return j;
    }
    return -1;
}

}  // namespace yue
