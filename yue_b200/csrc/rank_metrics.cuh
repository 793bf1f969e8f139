// K6: ranking metrics on the device (replaces Measure.rankingMeasure, evaluation/measure.py:16-41,
// for the lists of the last yue_rank_topn call -- SURVEY.md 8(f) row 2: with 1 M test users the
// reference's Python set operations, not the ranking, dominate evalRanking).
//
// Per user row and per cut-off n (the -topN list, item.ranking in the config):
//   hits       |test(u) ∩ pred(u)[:n]|                                   measure.py:7-13
//   recall     hits / |test(u)|                                          measure.py:91-94
//   AP         sum over hit ranks r of (hits so far)/(r+1), / min(|test(u)|, n)   measure.py:56-66
//   NDCG       sum_hit 1/log2(r+2)  /  sum_{r < min(|test(u)|, n)} 1/log2(r+2)    (the reference has no NDCG)
// and over all rows the number of distinct recommended tracks (coverage, measure.py:43-48).
// One warp per row: the hit pattern of the (<= 128) ranks is four ballot words, everything else is
// popcounts on them.  Per-row terms are written out and reduced by a second kernel in a FIXED order,
// so the sums do not depend on scheduling; hits are integers, the others float64.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "philox.cuh"   // row_contains

namespace yue {

constexpr int kMaxCuts = 8;

struct MetricsParams {
    const int32_t* ids;          // [B, N] lists of the last ranking call, -1 padded
    const int32_t* users;        // [B] local user of each row
    int64_t B;
    int N;
    const int64_t* test_indptr;  // [m+1]
    const int32_t* test_items;   // sorted unique held-out tracks per user
    int n_cuts;
    int cuts[kMaxCuts];
    double* terms;               // [n_cuts][4][B]: hits, recall term, AP, NDCG
    uint32_t* seen;              // [n_cuts][ceil(n_items/32)] bitmap of recommended tracks
    int64_t seen_words;
};

__global__ void __launch_bounds__(256) rank_metrics_kernel(const MetricsParams p) {
    __shared__ double disc[129];                 // disc[r] = sum_{x<r} 1/log2(x+2)
    if (threadIdx.x == 0) {
        double acc = 0.0;
        disc[0] = 0.0;
        for (int r = 0; r < 128; ++r) { acc += 1.0 / log2((double)r + 2.0); disc[r + 1] = acc; }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= p.B) return;
    const int u = p.users[b];
    const int64_t t0 = p.test_indptr[u];
    const int tlen = (int)(p.test_indptr[u + 1] - t0);
    uint32_t hit[4] = {0u, 0u, 0u, 0u};
    int32_t my_id[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int r = 32 * c + lane;
        my_id[c] = r < p.N ? p.ids[b * p.N + r] : -1;
        const bool h = my_id[c] >= 0 && row_contains(p.test_items + t0, tlen, my_id[c]);
        hit[c] = __ballot_sync(0xffffffffu, h);
    }
    for (int k = 0; k < p.n_cuts; ++k) {
        const int n = p.cuts[k];
        int hits = 0;
        double ap = 0.0, dcg = 0.0;
        int before = 0;                              // hits in earlier words
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int r = 32 * c + lane;
            const uint32_t in_cut = n - 32 * c >= 32 ? 0xffffffffu : (n - 32 * c <= 0 ? 0u : ((1u << (n - 32 * c)) - 1u));
            const uint32_t w = hit[c] & in_cut;
            if ((w >> lane) & 1u) {
                const int found = before + __popc(w & (0xffffffffu >> (31 - lane)));
                ap += (double)found / (double)(r + 1);
                dcg += 1.0 / log2((double)r + 2.0);
            }
            if (r < n && my_id[c] >= 0) atomicOr(p.seen + (size_t)k * p.seen_words + (my_id[c] >> 5), 1u << (my_id[c] & 31));
            hits += __popc(w);
            before += __popc(w);
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            ap += __shfl_xor_sync(0xffffffffu, ap, m);
            dcg += __shfl_xor_sync(0xffffffffu, dcg, m);
        }
        if (lane == 0) {
            const int lim = tlen < n ? tlen : n;
            double* t = p.terms + (size_t)k * 4 * p.B;
            t[b] = (double)hits;
            t[p.B + b] = tlen > 0 ? (double)hits / (double)tlen : 0.0;
            t[2 * p.B + b] = lim > 0 ? ap / (double)lim : 0.0;
            t[3 * p.B + b] = lim > 0 ? dcg / disc[lim < 128 ? lim : 128] : 0.0;
        }
    }
}

// deterministic sum of each [B] term vector: one block per vector, fixed strides, fixed tree
__global__ void __launch_bounds__(1024) metrics_reduce_kernel(const double* __restrict__ terms, int64_t B, double* __restrict__ out) {
    __shared__ double sm[1024];
    const double* t = terms + (size_t)blockIdx.x * B;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < B; i += 1024) acc += t[i];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 512; s >= 1; s >>= 1) {
        if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}

__global__ void popcount_kernel(const uint32_t* __restrict__ words, int64_t n_words, unsigned long long* __restrict__ out) {
    unsigned long long acc = 0;
    const uint32_t* w = words + (size_t)blockIdx.y * n_words;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_words; i += (int64_t)gridDim.x * blockDim.x) acc += __popc(w[i]);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out + blockIdx.y, acc);      // integer: order-free
}

}  // namespace yue
