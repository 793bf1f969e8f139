// K0: play log -> array form on the device.
//
// Replaces the array-building half of Record.preprocess (data/record.py:138-202) and the userListen /
// iteration-order build of recommender/cf/BPR.py:32-45 for logs whose entities are already numbered
// (ids by first appearance stay with the host, data/record.py:138-146): from the events in FILE ORDER
//     ev_user[E], ev_item[E], is_test[E]
// it builds, without leaving the GPU,
//   * ev_indptr / ev_items   training events grouped by user, file order kept inside a user, repeat
//                            plays kept (BPR.py:42-45)                      -> stable radix sort by user
//   * uq_indptr / uq_items   per-user sorted unique played tracks (BPR.py:32-35) -> sort + unique of
//                            64-bit keys user << 32 | track
//   * test_indptr / test_items  Record.testSet: unique held-out (user, track) pairs MINUS the pairs the
//                            user has in training (record.py:195-202)      -> sort + unique + anti-join
// The reference does this with dict-of-dict inserts, ~1 us of interpreter per event; at config C2
// (62.5 M events) numpy's argsort/unique take ~10 s on the host.  CUB's radix sort and select are the
// only library pieces (plain device-wide primitives, like calling cuBLAS for a plain GEMM).
#pragma once
#include <cstdint>
#include <cub/cub.cuh>
#include <cuda_runtime.h>

namespace yue {

__global__ void ingest_keys_kernel(const int32_t* __restrict__ u, const int32_t* __restrict__ it, int64_t E,
                                   uint64_t* __restrict__ keys) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x)
        keys[e] = ((uint64_t)(uint32_t)u[e] << 32) | (uint32_t)it[e];
}
__global__ void ingest_flags_kernel(const uint8_t* __restrict__ is_test, int64_t E, uint8_t* __restrict__ train_flag,
                                    uint8_t* __restrict__ test_flag) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t t = is_test ? (is_test[e] != 0) : 0;
        train_flag[e] = !t; test_flag[e] = t;
    }
}
// indptr[u] = first position whose user (high word of a sorted key array, or a sorted user array) is >= u
__global__ void ingest_indptr_from_keys_kernel(const uint64_t* __restrict__ keys, int64_t count, int64_t m, int64_t* __restrict__ indptr) {
    for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u <= m; u += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = count;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if ((int64_t)(keys[mid] >> 32) < u) lo = mid + 1; else hi = mid; }
        indptr[u] = lo;
    }
}
__global__ void ingest_indptr_from_users_kernel(const int32_t* __restrict__ users, int64_t count, int64_t m, int64_t* __restrict__ indptr) {
    for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u <= m; u += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = count;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if ((int64_t)users[mid] < u) lo = mid + 1; else hi = mid; }
        indptr[u] = lo;
    }
}
__global__ void ingest_low_words_kernel(const uint64_t* __restrict__ keys, int64_t count, int32_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (int32_t)(uint32_t)keys[i];
}
// flag[i] = 1 when test key i is NOT among the (sorted unique) training keys
__global__ void ingest_antijoin_kernel(const uint64_t* __restrict__ test_keys, int64_t nt, const uint64_t* __restrict__ train_keys,
                                       int64_t ntr, uint8_t* __restrict__ flag) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nt; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = test_keys[i];
        int64_t lo = 0, hi = ntr;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (train_keys[mid] < k) lo = mid + 1; else hi = mid; }
        flag[i] = !(lo < ntr && train_keys[lo] == k);
    }
}
__global__ void ingest_range_check_kernel(const int32_t* __restrict__ u, const int32_t* __restrict__ it, int64_t E, int64_t m,
                                          int64_t n, int* __restrict__ bad) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x)
        if (u[e] < 0 || u[e] >= m || it[e] < 0 || it[e] >= n) *bad = 1;
}

// Arrays handed to yue_set_interactions: every track id in [0, n) and every play row strictly increasing (the rejection
// sampler and the ranking mask search the rows; an id outside the catalog would index Q and the counters out of bounds).
// bad[0] |= 1: event id out of range, 2: play-row id out of range, 4: play row not sorted-unique.
__global__ void interaction_check_kernel(const int32_t* __restrict__ ev_items, int64_t T, const int64_t* __restrict__ uq_indptr,
                                         const int32_t* __restrict__ uq_items, int64_t nnz, int64_t m, int64_t n, int* __restrict__ bad) {
    int flags = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int64_t e = t0; e < T; e += stride) if (ev_items[e] < 0 || ev_items[e] >= n) flags |= 1;
    for (int64_t x = t0; x < nnz; x += stride) if (uq_items[x] < 0 || uq_items[x] >= n) flags |= 2;
    // row starts are the only places where uq_items[x - 1] >= uq_items[x] is allowed: clear those, then every other descent is an error
    for (int64_t x = t0 + 1; x < nnz; x += stride) {
        if (uq_items[x - 1] >= uq_items[x]) {
            int64_t lo = 0, hi = m;                                  // is x the first entry of a row?  (rare path: binary search)
            while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (uq_indptr[mid] < x) lo = mid + 1; else hi = mid; }
            if (!(lo <= m && uq_indptr[lo] == x)) flags |= 4;
        }
    }
    if (flags) atomicOr(bad, flags);
}

}  // namespace yue
