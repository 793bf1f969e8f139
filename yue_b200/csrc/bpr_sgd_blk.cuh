// K1+K2, blocked form: the throughput-mode BPR update for rows of up to 128 floats
// (replaces recommender/cf/BPR.py:42-58, same schedule and sampler as bpr_sgd.cuh).
// Rows of exactly 32, 64 or 128 floats use every lane (MASK = false); any other width (num.factors = 10, 20, 50, 100 ...:
// config/BPR.conf ships 10) runs the same text with the lanes past the row's end switched off (MASK = true): their
// registers stay zero, so they add nothing to a dot product, and they neither load nor add.
//
// Why a second form.  ncu on bpr_sgd_kernel (profiles/ncu_summary_r1.md): 251 warp instructions per
// triplet, issued one every ~7 cycles per warp -- the kernel is bound by the DEPENDENT instruction
// chain of one warp, not by memory: consecutive triplets of a user are serial through P[u]
// (x_t = P[u]_t . (Q[i_t] - Q[j_t]) needs the P[u] the previous triplet wrote), and every x_t costs a
// 6-stage shuffle reduction.  This kernel removes the chain without changing the arithmetic:
//
//   for a block of K = 4 consecutive triplets of one user, with d_a = Q[i_a] - Q[j_a] (rows as read
//   at the start of the block) and c = 1 - lr*regU, the reference's update P <- c (P + g_a d_a) gives
//       x_a = P_a . d_a,     P_{a+1} . d_b = c (P_a . d_b + g_a  d_a . d_b)        (b > a)
//   so the four scores follow from the 4 dots  P_0 . d_a  and the 6 dots  d_a . d_b  by a scalar
//   recurrence.  The 10 dots are independent: one reduce-scatter/all-gather over the warp (27
//   shuffles per block instead of 40 dependent ones), then sigmoid/gradient for the four triplets
//   in scalar code, then the row updates (which need no further reductions).
//
// Rows of one block are read before any of its updates is published -- the same staleness the
// prefetching kernel has (it reads rows 4 triplets ahead), so a track that repeats inside a block
// is read at its pre-block value.
//
// Row layout across the warp: lane l owns floats [V*l, V*l + V) of EVERY row (P[u], Q[i], Q[j]),
// V = ld / 32, so d_a and all partial dots are lane-local and a row moves as one coalesced
// 128/256/512-byte request (LDG.32/64/128, RED.ADD.F32 / .v2 / .v4).
//
// Hot rows: see "hot-row table" below.
#pragma once
#include "bpr_sgd.cuh"

namespace yue {

constexpr int kBlkThreads = 384;      // <= 12 warps per CTA: up to 168 registers per thread
constexpr int kBlkK = 4;              // triplets per block

template <int V> __device__ __forceinline__ void ldv(const float* p, float (&v)[V]);
template <> __device__ __forceinline__ void ldv<1>(const float* p, float (&v)[1]) { v[0] = __ldcg(p); }
template <> __device__ __forceinline__ void ldv<2>(const float* p, float (&v)[2]) {
    const float2 t = __ldcg(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1] = t.y;
}
template <> __device__ __forceinline__ void ldv<4>(const float* p, float (&v)[4]) {
    const float4 t = __ldcg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <int V> __device__ __forceinline__ void redv(float* p, const float (&v)[V]);
template <> __device__ __forceinline__ void redv<1>(float* p, const float (&v)[1]) {
    asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" :: "l"(p), "f"(v[0]) : "memory");
}
template <> __device__ __forceinline__ void redv<2>(float* p, const float (&v)[2]) {
    asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" :: "l"(p), "f"(v[0]), "f"(v[1]) : "memory");
}
template <> __device__ __forceinline__ void redv<4>(float* p, const float (&v)[4]) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}

// The same at system scope, for rows that may live in a peer GPU's memory (yue_hot_share): a strong load is served by
// the L2 of the GPU that owns the line, the add is performed there; relaxed -- the schedule is Hogwild.
template <int V> __device__ __forceinline__ void ldv_sys(const float* p, float (&v)[V]);
template <> __device__ __forceinline__ void ldv_sys<1>(const float* p, float (&v)[1]) {
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v[0]) : "l"(p) : "memory");
}
template <> __device__ __forceinline__ void ldv_sys<2>(const float* p, float (&v)[2]) {
    asm volatile("ld.relaxed.sys.global.v2.f32 {%0, %1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "l"(p) : "memory");
}
template <> __device__ __forceinline__ void ldv_sys<4>(const float* p, float (&v)[4]) {
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p) : "memory");
}
template <int V> __device__ __forceinline__ void redv_sys(float* p, const float (&v)[V]);
template <> __device__ __forceinline__ void redv_sys<1>(float* p, const float (&v)[1]) {
    asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" :: "l"(p), "f"(v[0]) : "memory");
}
template <> __device__ __forceinline__ void redv_sys<2>(float* p, const float (&v)[2]) {
    asm volatile("red.relaxed.sys.global.add.v2.f32 [%0], {%1, %2};" :: "l"(p), "f"(v[0]), "f"(v[1]) : "memory");
}
template <> __device__ __forceinline__ void redv_sys<4>(float* p, const float (&v)[4]) {
    asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}

// Sum each of 10 per-lane values over the warp; every lane gets all 10 totals.  Values 0..7 go
// through a reduce-scatter (xor 16, 8, 4 halve the live set) + butterfly (xor 2, 1) + all-gather,
// values 8..9 through a plain butterfly: 27 shuffles instead of 50.
__device__ __forceinline__ void warp_allreduce10(float (&x)[10], int lane) {
    const unsigned full = 0xffffffffu;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float k0[4], k1[2], k2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b4 ? x[i] : x[i + 4], keep = b4 ? x[i + 4] : x[i];
        k0[i] = keep + __shfl_xor_sync(full, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b3 ? k0[i] : k0[i + 2], keep = b3 ? k0[i + 2] : k0[i];
        k1[i] = keep + __shfl_xor_sync(full, send, 8);
    }
    {
        const float send = b2 ? k1[0] : k1[1], keep = b2 ? k1[1] : k1[0];
        k2 = keep + __shfl_xor_sync(full, send, 4);
    }
    float y8 = x[8], y9 = x[9];
    y8 += __shfl_xor_sync(full, y8, 16); y9 += __shfl_xor_sync(full, y9, 16);
    y8 += __shfl_xor_sync(full, y8, 8);  y9 += __shfl_xor_sync(full, y9, 8);
    y8 += __shfl_xor_sync(full, y8, 4);  y9 += __shfl_xor_sync(full, y9, 4);
    k2 += __shfl_xor_sync(full, k2, 2);  y8 += __shfl_xor_sync(full, y8, 2); y9 += __shfl_xor_sync(full, y9, 2);
    k2 += __shfl_xor_sync(full, k2, 1);  y8 += __shfl_xor_sync(full, y8, 1); y9 += __shfl_xor_sync(full, y9, 1);
    // lane l now holds the total of value (l >> 2)
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __shfl_sync(full, k2, 4 * i);
    x[8] = y8; x[9] = y9;
}

// The same for 16 values: reduce-scatter over all five lane bits, all-gather (32 shuffles instead of 80).
__device__ __forceinline__ void warp_allreduce16(float (&x)[16], int lane) {
    const unsigned full = 0xffffffffu;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
    float k0[8], k1[4], k2[2], k3;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = b4 ? x[i] : x[i + 8], keep = b4 ? x[i + 8] : x[i];
        k0[i] = keep + __shfl_xor_sync(full, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b3 ? k0[i] : k0[i + 4], keep = b3 ? k0[i + 4] : k0[i];
        k1[i] = keep + __shfl_xor_sync(full, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b2 ? k1[i] : k1[i + 2], keep = b2 ? k1[i + 2] : k1[i];
        k2[i] = keep + __shfl_xor_sync(full, send, 4);
    }
    {
        const float send = b1 ? k2[0] : k2[1], keep = b1 ? k2[1] : k2[0];
        k3 = keep + __shfl_xor_sync(full, send, 2);
    }
    k3 += __shfl_xor_sync(full, k3, 1);
    // lane l now holds the total of value (l >> 1)
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = __shfl_sync(full, k3, 2 * i);
}

// g = lr * sigmoid(-x) (= lr (1 - s), BPR.py:50-51) and loss += -log(sigmoid(x)) (BPR.py:58)
__device__ __forceinline__ float bpr_grad(float x, float lr, float& loss) {
    const float ex = __expf(-fabsf(x));                     // e^{-|x|} in (0, 1]
    loss += fmaxf(-x, 0.f) + __logf(1.f + ex);
    return lr * __fdividef(x >= 0.f ? ex : 1.f, 1.f + ex);
}

template <int V>
struct BlkRows {
    float qi[kBlkK][V], qj[kBlkK][V];
    float* pi[kBlkK];                // this lane's part of the positive / negative row of each triplet
    float* pj[kBlkK];                // (in Q, or in the hot-row table)
    float ex[kBlkK][V];              // second accumulator row of a sharded hot positive (valid when dx != 0)
    int dx[kBlkK];                   // byte distance from pi to that row, 0 = the positive is not sharded
};



// ---- hot-row table ---------------------------------------------------------------------------
// Measured on B200 (tools/red_probe*.cu, profiles/red_probe_r1.md): an L2 slice serves about one
// 32-byte sector per clock, loads and atomics alike, a 256-byte row stored contiguously lives in ONE
// slice (the slice hash ignores address bits 0-7 and 9), so one load + one RED of the most played
// track costs its slice 8.4 ns -- 33 ms per epoch at config C2, where that track is the positive of
// 7.8 % of all triplets; ncu showed that slice's atomic unit 86-90 % busy and the others 5-7 %.
// Accumulator shards (bpr_sgd.cuh) only divide the REDs: every reader still reads every shard.
// Here the rows of the hot tracks live, for the duration of a launch, in a small table whose layout
// puts the SECTORS of a row into different slices: sector c of slot s sits at
//     c * kHotPlane + granule(s) * 256 B,   granule(s) = (s / 2) * 4 + s % 2   (bit 9 of the address stays 0)
// so a touch of the row costs each of 8 slices one load sector and one atomic sector: 2.2 ns per touch
// of a single row, 0.9-1.2 ns per touch over a zipf mix of rows.  hot_gather/hot_scatter kernels copy
// the rows in from Q before the launch and back after it.
// The most played tracks of all additionally get a SECOND accumulator row in the table (the logical
// row is the sum of both; a warp adds to row warp % 2, readers fetch both): one 32-byte address takes
// a dependent load + atomic every ~2.2 ns, which for the top track of C2 is still 8.6 ms per epoch.
constexpr int kHotSlots = 248;                                // hot tracks: one GPU uses ~10 (tracks above 1/128 of the events); the sharded
                                                              // trainer shares up to all of them between the ranks (tracks above 1/4096)
constexpr int kHotExtra = 8;                                  // of which so many may have a second row
constexpr int kHotRows = kHotSlots + kHotExtra;
// floats between two sectors of a row ("plane"): 2 x rows + 1 granules of 256 B.  65 granules for tables of up to 32 rows (the
// stride measured in tools/red_probe3.cu, on which the one-GPU numbers rest), 513 for larger ones.
constexpr int kHotPlaneSmall = (2 * 32 + 1) * 64, kHotPlaneBig = (2 * kHotRows + 1) * 64;
__host__ __device__ __forceinline__ int hot_plane_floats(int rows) { return rows <= 32 ? kHotPlaneSmall : kHotPlaneBig; }
__host__ __device__ __forceinline__ size_t hot_slot_offset(int row) {       // floats; row = slot, or n_hot + k for an extra row
    return (size_t)(((row >> 1) << 2) | (row & 1)) * 64;
}
// float offset, inside a slot's row, of the V floats lane `lane` owns
template <int V> __host__ __device__ __forceinline__ size_t hot_lane_offset(int lane, int plane) {
    return (size_t)((V * lane) >> 3) * plane + ((V * lane) & 7);
}
constexpr size_t kHotTableFloats = (size_t)16 * kHotPlaneBig;              // up to 16 sectors (ld = 128)
constexpr int kHotCandidates = 8;        // placements of the table the first launch on a log times (yue_b200.cu)
constexpr int kHotCandidateStep = 11;    // granules between them

template <int V>
__global__ void hot_gather_kernel(const float* __restrict__ Q, float* __restrict__ hotQ, const int32_t* __restrict__ hot_items,
                                  const int32_t* __restrict__ hot_dx, int n_hot, int ld, int plane) {
    const int lane = threadIdx.x & 31;
    for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_hot; s += (gridDim.x * blockDim.x) >> 5) {
        float* row = hotQ + hot_slot_offset(s) + hot_lane_offset<V>(lane, plane);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            row[v] = V * lane + v < ld ? Q[(size_t)hot_items[s] * ld + V * lane + v] : 0.f;
            if (hot_dx[s]) row[hot_dx[s] / 4 + v] = 0.f;
        }
    }
}
template <int V>
__global__ void hot_scatter_kernel(float* __restrict__ Q, const float* __restrict__ hotQ, const int32_t* __restrict__ hot_items,
                                   const int32_t* __restrict__ hot_dx, int n_hot, int ld, int plane) {
    const int lane = threadIdx.x & 31;
    for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_hot; s += (gridDim.x * blockDim.x) >> 5) {
        const float* row = hotQ + hot_slot_offset(s) + hot_lane_offset<V>(lane, plane);
#pragma unroll
        for (int v = 0; v < V; ++v)
            if (V * lane + v < ld) Q[(size_t)hot_items[s] * ld + V * lane + v] = row[v] + (hot_dx[s] ? row[hot_dx[s] / 4 + v] : 0.f);
    }
}

// yue_hot_share: slot s lives in the table of rank s % nranks (hot_base[s] = this process's mapping of it).
// gather: the slots THIS rank owns move from its Q into its own table (second rows zeroed).
template <int V>
__global__ void hot_gather_shared_kernel(const float* __restrict__ Q, const unsigned long long* __restrict__ hot_base,
                                         const int32_t* __restrict__ hot_items, const int32_t* __restrict__ hot_dx, int n_hot, int ld,
                                         int nranks, int rank, int plane) {
    const int lane = threadIdx.x & 31;
    for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_hot; s += (gridDim.x * blockDim.x) >> 5) {
        if (s % nranks != rank) continue;
        float* row = reinterpret_cast<float*>(hot_base[s]) + hot_slot_offset(s) + hot_lane_offset<V>(lane, plane);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            row[v] = Q[(size_t)hot_items[s] * ld + V * lane + v];
            if (hot_dx[s]) row[hot_dx[s] / 4 + v] = 0.f;
        }
    }
}
// pull: every hot row (the sum of its accumulator rows) from its owner's table into this rank's Q
template <int V>
__global__ void hot_pull_shared_kernel(float* __restrict__ Q, const unsigned long long* __restrict__ hot_base,
                                       const int32_t* __restrict__ hot_items, const int32_t* __restrict__ hot_dx, int n_hot, int ld, int plane) {
    const int lane = threadIdx.x & 31;
    for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_hot; s += (gridDim.x * blockDim.x) >> 5) {
        const float* row = reinterpret_cast<const float*>(hot_base[s]) + hot_slot_offset(s) + hot_lane_offset<V>(lane, plane);
        float a[V], b[V];
        ldv_sys<V>(row, a);
        if (hot_dx[s]) ldv_sys<V>(row + hot_dx[s] / 4, b);
#pragma unroll
        for (int v = 0; v < V; ++v) Q[(size_t)hot_items[s] * ld + V * lane + v] = a[v] + (hot_dx[s] ? b[v] : 0.f);
    }
}

// Control flow.  A warp's work is a stream of segments (items from the global cursor, segments of an
// item consecutive).  Per segment there is ONE loop over its blocks of 4 triplets; the pass over the
// last block first prepares the NEXT segment -- draws its 32 negatives (K1) and issues the loads of its
// first block -- so those loads fly while the last block is computed, and at the same point the
// segment after that is prefetched (record, positives, play row, P[u]).  An item starts with a
// virtual empty segment, so all of this exists once in the code (the first version, with the block
// loop unrolled by two and the prologue inlined three times, was 12 K instructions and spent 57 % of
// its stall samples waiting for the instruction cache).
// SHARED (yue_hot_share, N GPUs): slot s of the hot-row table lives in the table of rank s % N -- p.hot_base[s] is this
// process's mapping of that table -- and every access to a row of Q or of a table is made at system scope.
template <int V, bool APR, bool MASK = false, bool SHARED = false>
__global__ void __launch_bounds__(kBlkThreads, 1) bpr_sgd_blk_kernel(const SgdParams p) {
    static_assert(!(SHARED && MASK), "shared hot rows need full-width rows");
    extern __shared__ __align__(16) int hot_sm[];           // [n_hot] hot track ids, ascending; [n_hot] their slots; [n_hot] dx
    int* hot_sorted = hot_sm;
    int* hot_sorted_slot = hot_sm + p.n_hot;
    int* hot_dx = hot_sm + 2 * p.n_hot;                     // per slot: byte distance to its second row, or 0
    unsigned long long* hot_base = reinterpret_cast<unsigned long long*>(hot_sm + ((3 * p.n_hot + 3) & ~3));   // SHARED: [n_hot]
    for (int x = threadIdx.x; x < p.n_hot; x += blockDim.x) {
        hot_sorted[x] = p.hot_sorted[x]; hot_sorted_slot[x] = p.hot_sorted_slot[x]; hot_dx[x] = p.hot_dx[x];
        if (SHARED) hot_base[x] = p.hot_base[x];
    }
    __syncthreads();
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warp = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (warp >= p.n_warps) return;
    const int lane_off = V * lane;
    const bool act = !MASK || lane_off < p.ld;               // does this lane hold a part of the row?  (ld % 4 == 0, V | 4)
    auto ldrow = [&](const float* ptr, float (&v)[V]) {
        if (!MASK || act) ldv<V>(ptr, v);
        else {
#pragma unroll
            for (int x = 0; x < V; ++x) v[x] = 0.f;
        }
    };
    auto redrow = [&](float* ptr, const float (&v)[V]) { if (!MASK || act) redv<V>(ptr, v); };
    // rows of Q / of a hot-row table
    auto ldq = [&](const float* ptr, float (&v)[V]) { if (SHARED) ldv_sys<V>(ptr, v); else ldrow(ptr, v); };
    auto redq = [&](float* ptr, const float (&v)[V]) { if (SHARED) redv_sys<V>(ptr, v); else redrow(ptr, v); };
    const float cu1 = 1.f - p.c_u;
    // per-lane base addresses; a row is then base + (32-bit byte offset), see q_ptr
    char* const q_lane = reinterpret_cast<char*>(p.Q + lane_off);
    char* const hot_lane = reinterpret_cast<char*>(p.hotQ + hot_lane_offset<V>(lane, p.hot_plane));
    const uint32_t hot_lane_bytes = (uint32_t)(hot_lane_offset<V>(lane, p.hot_plane) * sizeof(float));
    const uint32_t row_bytes = (uint32_t)p.ld * 4u;          // n * ld * 4 < 4 GB is checked by the host

    float pu[V], pu0[V], pun[V];
#pragma unroll
    for (int v = 0; v < V; ++v) pu[v] = pu0[v] = pun[v] = 0.f;
    int cur_u = -1;
    double loss = 0.0;

    // address of this lane's part of track row `t` (t < 0: hot slot -t-1)
    auto q_ptr = [&](int32_t t) -> float* {
        const int s = -t - 1;
        const uint32_t off = t < 0 ? (uint32_t)(((s >> 1) << 2) | (s & 1)) * 256u : (uint32_t)t * row_bytes;
        if (SHARED && t < 0) return reinterpret_cast<float*>(hot_base[s] + hot_lane_bytes + off);
        return reinterpret_cast<float*>((t < 0 ? hot_lane : q_lane) + off);
    };
    auto flush_user = [&]() {
        if (cur_u < 0) return;
        float dlt[V];
#pragma unroll
        for (int v = 0; v < V; ++v) dlt[v] = pu[v] - pu0[v];
        redrow(p.P + (size_t)cur_u * p.ld + lane_off, dlt);
    };
    auto sync_user = [&](int uu) {          // publish the pending change of P[cur_u], (re)load P[uu]
        flush_user();
        cur_u = uu;
        ldrow(p.P + (size_t)uu * p.ld + lane_off, pu);
#pragma unroll
        for (int v = 0; v < V; ++v) pu0[v] = pu[v];
    };
    auto take_item = [&]() -> int64_t {
        unsigned long long it = 0;
        if (lane == 0) it = atomicAdd(p.cursor, 1ull);
        return (int64_t)__shfl_sync(full, it, 0);
    };
    auto load_rec = [&](int64_t seg, int4& a, int4& b) {
        const int4* r = reinterpret_cast<const int4*>(p.seg_rec + seg);
        a = __ldg(r); b = __ldg(r + 1);
    };
    auto rec_begin = [](const int4& a) { return (int64_t)(((uint64_t)(uint32_t)a.w << 32) | (uint32_t)a.z); };
    auto rec_row = [](const int4& b) { return (int64_t)(((uint64_t)(uint32_t)b.y << 32) | (uint32_t)b.x); };
    // pull what the prologue of a segment will read into L1/L2 (no register results, nothing to wait for)
    auto prefetch_seg = [&](const int4& a, const int4& b) {
        const int64_t begin = rec_begin(a);
        if (lane == 0) {
            asm volatile("prefetch.global.L1 [%0];" :: "l"(p.ev_items + begin));
            if (p.ev_neg) asm volatile("prefetch.global.L1 [%0];" :: "l"(p.ev_neg + begin));
        }
        if (lane == 1) asm volatile("prefetch.global.L1 [%0];" :: "l"(p.ev_items + begin + ((a.y & 63) - 1)));
        if (8 * lane < b.z) asm volatile("prefetch.global.L1 [%0];" :: "l"(p.uq_items + rec_row(b) + 8 * lane));
        if (lane < (int)(row_bytes / 32u)) asm volatile("prefetch.global.L2 [%0];" :: "l"(p.P + (size_t)a.x * p.ld + 8 * lane));
    };

    int64_t item = take_item();
    while (item < p.n_work) {
        const int64_t next_item = take_item();
        const int64_t sb = p.item_ptr[2 * item], se = p.item_ptr[2 * item + 1];
        // record queue: (na, nb) = segment seg+1, (fa, fb) = segment seg+2
        int4 na, nb, fa = make_int4(0, 0, 0, 0), fb = fa;
        load_rec(sb, na, nb);
        if (sb + 1 < se) load_rec(sb + 1, fa, fb);
        // current segment: starts as a virtual empty one in front of segment sb
        int u = cur_u, len = 0;
        bool resync = false;
        int32_t my_i = 0, my_j = 0;
        // two row buffers that swap roles every block (the block being computed / the block whose loads are in flight):
        // the body below exists twice, once per assignment, selected by a warp-uniform bit -- a copy `cur = nxt` per
        // block was 44 MOVs, 9.4 of the 158 instructions per triplet (profiles/ncu_summary_r1.md)
        BlkRows<V> rowsA, rowsB;
        bool flip = false;                                   // false: cur = rowsA, nxt = rowsB
#pragma unroll
        for (int a = 0; a < kBlkK; ++a) {
            rowsA.pi[a] = rowsA.pj[a] = nullptr;
            rowsA.dx[a] = 0;
#pragma unroll
            for (int v = 0; v < V; ++v) rowsA.qi[a][v] = rowsA.qj[a][v] = rowsA.ex[a][v] = 0.f;
        }
        for (int64_t seg = sb - 1; seg < se; ++seg) {
            const int nblk = (len + kBlkK - 1) / kBlkK;
            const bool has_next = seg + 1 < se;
            int32_t n_i = 0, n_j = 0;
            int n_u = 0, n_len = 0;
            bool n_resync = false;
#pragma unroll 1
            for (int b = 0; b == 0 || b < nblk; ++b) {
                const bool last = b + 1 >= nblk;
                if (last && has_next) {
                    // ---- prepare segment seg+1 (record na/nb arrived; its data was prefetched) ----
                    n_u = na.x; n_len = na.y & 63; n_resync = nb.w != 0;
                    if (lane < n_len) {
                        const int64_t e = rec_begin(na) + lane;
                        n_i = p.ev_items[e];             // hot positives arrive re-labelled -slot-1
                        n_j = p.ev_neg ? p.ev_neg[e]    // K1: lane t draws the negative of event begin+t
                                       : sample_negative(p.seed, p.epoch, (uint64_t)((p.ev_delta ? p.ev_delta[n_u] : p.event_base) + e), p.slot,
                                                         p.n_items, p.uq_items + rec_row(nb), nb.z);
                        if (p.n_hot > 0) {               // a negative that happens to be a hot track lives in the table too
                            int lo = 0, hi = p.n_hot;
                            while (lo < hi) { const int mid = (lo + hi) >> 1; if (hot_sorted[mid] < n_j) lo = mid + 1; else hi = mid; }
                            if (lo < p.n_hot && hot_sorted[lo] == n_j) n_j = -hot_sorted_slot[lo] - 1;
                        }
                    }
                    if (p.pf_stride > 0 && lane < n_len) {
                        // Q beyond the L2 (config C3: 1 GB): the segment's 64 rows start their way from DRAM now, 1..8 blocks
                        // before the register loads that use them (those are issued one block ahead only)
                        const char* ri = reinterpret_cast<const char*>(p.Q) + (size_t)(n_i < 0 ? 0 : n_i) * row_bytes;
                        const char* rj = reinterpret_cast<const char*>(p.Q) + (size_t)(n_j < 0 ? 0 : n_j) * row_bytes;
                        for (uint32_t o = 0; o < row_bytes; o += (uint32_t)p.pf_stride) {
                            if (n_i >= 0) asm volatile("prefetch.global.L2 [%0];" :: "l"(ri + o));
                            if (n_j >= 0) asm volatile("prefetch.global.L2 [%0];" :: "l"(rj + o));
                        }
                    }
                    __syncwarp();
                    if (n_u != cur_u) ldrow(p.P + (size_t)n_u * p.ld + lane_off, pun);
                    // the segment after: its record arrived a segment ago -> prefetch its data, fetch the next record
                    if (seg + 2 < se) prefetch_seg(fa, fb);
                    na = fa; nb = fb;
                    if (seg + 3 < se) load_rec(seg + 3, fa, fb);
                }
                auto half = [&](BlkRows<V>& cur, BlkRows<V>& nxt) {
                    // ---- loads of the next block: block b+1 of this segment, or block 0 of the next one ----
                    if (!last || has_next) {
                        const int32_t src_i = last ? n_i : my_i, src_j = last ? n_j : my_j;
                        const int src_len = last ? n_len : len, t0 = last ? 0 : kBlkK * (b + 1);
#pragma unroll
                        for (int a = 0; a < kBlkK; ++a) {
                            const int t = t0 + a;                               // <= 31
                            const int32_t it = __shfl_sync(full, src_i, t), jt = __shfl_sync(full, src_j, t);
                            nxt.dx[a] = 0;
                            if (t < src_len) {
                                nxt.pi[a] = q_ptr(it); nxt.pj[a] = q_ptr(jt);
                                ldq(nxt.pi[a], nxt.qi[a]);
                                ldq(nxt.pj[a], nxt.qj[a]);
                                // A sharded hot track drawn as the NEGATIVE is read through its first accumulator row only (its
                                // second row would need V more registers per triplet in a kernel that has none to spare): the
                                // score then misses the other row's share of the launch's changes.  A negative is a uniform draw
                                // over the catalog, so at most kHotExtra / n of the triplets (config C2: 8 / 200 000) are
                                // affected, and only in the value read, never in what is added.
                                if (it < 0) {                                   // a sharded hot positive: fetch its second row too
                                    nxt.dx[a] = hot_dx[-it - 1];
                                    if (nxt.dx[a]) ldq(reinterpret_cast<float*>(reinterpret_cast<char*>(nxt.pi[a]) + nxt.dx[a]), nxt.ex[a]);
                                }
                            } else {
#pragma unroll
                                for (int v = 0; v < V; ++v) nxt.qi[a][v] = nxt.qj[a][v] = 0.f;
                            }
                        }
                    }
                    if (b < nblk) {
                        // ---- K2: block b of the current segment ----------------------------------------
                        const int nbk = len - kBlkK * b;                        // >= 1 triplets in this block
                        if (resync && b > 0 && (b & p.resync_mask) == 0) sync_user(u);
                        float d[kBlkK][V];
#pragma unroll
                        for (int a = 0; a < kBlkK; ++a) {
                            if (cur.dx[a]) {                                    // logical row of a sharded positive = sum of its rows
#pragma unroll
                                for (int v = 0; v < V; ++v) cur.qi[a][v] += cur.ex[a][v];
                            }
#pragma unroll
                            for (int v = 0; v < V; ++v) d[a][v] = cur.qi[a][v] - cur.qj[a][v];   // Q[i] - Q[j], old rows (BPR.py:51)
                        }
                        if constexpr (!APR) {
                            float x[10];
#pragma unroll
                            for (int a = 0; a < kBlkK; ++a) {
                                float s = 0.f;
#pragma unroll
                                for (int v = 0; v < V; ++v) s = fmaf(pu[v], d[a][v], s);
                                x[a] = s;
                            }
                            {
                                int g = 4;
#pragma unroll
                                for (int a = 0; a < kBlkK; ++a)
#pragma unroll
                                    for (int c = a + 1; c < kBlkK; ++c) {
                                        float s = 0.f;
#pragma unroll
                                        for (int v = 0; v < V; ++v) s = fmaf(d[a][v], d[c][v], s);
                                        x[g++] = s;                              // order: 01 02 03 12 13 23
                                    }
                            }
                            warp_allreduce10(x, lane);
                            // scalar recurrence: w_c = P_a . d_c for the current a
                            float g[kBlkK];
                            float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
                            g[0] = bpr_grad(x[0], p.lr, l0);
                            float w1 = cu1 * fmaf(g[0], x[4], x[1]);
                            float w2 = cu1 * fmaf(g[0], x[5], x[2]);
                            float w3 = cu1 * fmaf(g[0], x[6], x[3]);
                            g[1] = bpr_grad(w1, p.lr, l1);
                            w2 = cu1 * fmaf(g[1], x[7], w2);
                            w3 = cu1 * fmaf(g[1], x[8], w3);
                            g[2] = bpr_grad(w2, p.lr, l2);
                            w3 = cu1 * fmaf(g[2], x[9], w3);
                            g[3] = bpr_grad(w3, p.lr, l3);
                            loss += (double)(l0 + (nbk > 1 ? l1 : 0.f) + (nbk > 2 ? l2 : 0.f) + (nbk > 3 ? l3 : 0.f));
                            // row updates in the reference's order (BPR.py:51-57): P first, Q with the updated P,
                            // then the three multiplicative shrinks; Q changes leave as deltas (vector atomics)
#pragma unroll
                            for (int a = 0; a < kBlkK; ++a) {
                                if (a < nbk) {
                                    const float ga = g[a];
                                    float di[V], dj[V];
#pragma unroll
                                    for (int v = 0; v < V; ++v) {
                                        const float pn = fmaf(ga, d[a][v], pu[v]);
                                        const float gp = ga * pn;
                                        di[v] = fmaf(-p.c_i, cur.qi[a][v] + gp, gp);    // (q + g p)(1 - c) - q
                                        dj[v] = fmaf(-p.c_i, cur.qj[a][v] - gp, -gp);
                                        pu[v] = fmaf(-p.c_u, pn, pn);
                                    }
                                    redq(reinterpret_cast<float*>(reinterpret_cast<char*>(cur.pi[a]) + ((warp & 1) ? cur.dx[a] : 0)), di);
                                    redq(cur.pj[a], dj);
                                }
                            }
                        } else {
                            // K2a, APR (recommender/advanced/APR.py:25-76, oracle/apr_ref.py): per triplet, from the OLD rows,
                            //   y_adv = y - 2 eps |P| - eps |d| + 2 eps^2 y / (|P||d|),  a = lr (s0 + regA s1),  b = lr regA s1 eps,
                            //   P <- c (alpha P + a d),  alpha = 1 - 2 b / |P|;   Q[i] += a P - (b/|d|) d,  Q[j] -= the same
                            // The block needs |d_a|^2 and |P_0|^2 besides the 10 dots of BPR (16 reduced values); then
                            //   P_{a+1}.d_c = c (alpha w_c + a G_ac),   |P_{a+1}|^2 = c^2 (alpha^2 |P_a|^2 + 2 alpha a w_a + a^2 G_aa).
                            float x[16];
                            float pp = 0.f;
#pragma unroll
                            for (int v = 0; v < V; ++v) pp = fmaf(pu[v], pu[v], pp);
#pragma unroll
                            for (int a = 0; a < kBlkK; ++a) {
                                float s = 0.f, q = 0.f;
#pragma unroll
                                for (int v = 0; v < V; ++v) { s = fmaf(pu[v], d[a][v], s); q = fmaf(d[a][v], d[a][v], q); }
                                x[a] = s; x[10 + a] = q;
                            }
                            {
                                int g = 4;
#pragma unroll
                                for (int a = 0; a < kBlkK; ++a)
#pragma unroll
                                    for (int c = a + 1; c < kBlkK; ++c) {
                                        float s = 0.f;
#pragma unroll
                                        for (int v = 0; v < V; ++v) s = fmaf(d[a][v], d[c][v], s);
                                        x[g++] = s;
                                    }
                            }
                            x[14] = pp; x[15] = 0.f;
                            warp_allreduce16(x, lane);
                            float w[kBlkK] = {x[0], x[1], x[2], x[3]};
                            const float G[kBlkK][kBlkK] = {{x[10], x[4], x[5], x[6]}, {x[4], x[11], x[7], x[8]},
                                                           {x[5], x[7], x[12], x[9]}, {x[6], x[8], x[9], x[13]}};
                            float np2 = x[14], lsum = 0.f;
                            float ca[kBlkK], calpha[kBlkK], cbd[kBlkK];           // a, alpha, b/|d| of each triplet
#pragma unroll
                            for (int a = 0; a < kBlkK; ++a) {
                                const float y = w[a];
                                const float inp = np2 > 0.f ? rsqrtf(np2) : 0.f, ind = G[a][a] > 0.f ? rsqrtf(G[a][a]) : 0.f;
                                const float n_p = np2 * inp, n_d = G[a][a] * ind;
                                const float ya = y - 2.f * p.eps * n_p - p.eps * n_d + 2.f * p.eps * p.eps * y * inp * ind;
                                const float e0 = __expf(-fabsf(y)), e1 = __expf(-fabsf(ya));
                                const float s0 = __fdividef(y >= 0.f ? e0 : 1.f, 1.f + e0);      // sigmoid(-y)
                                const float s1 = __fdividef(ya >= 0.f ? e1 : 1.f, 1.f + e1);
                                if (a < nbk) lsum += fmaxf(-y, 0.f) + __logf(1.f + e0) + p.regA * (fmaxf(-ya, 0.f) + __logf(1.f + e1));
                                const float aa = p.lr * (s0 + p.regA * s1), bb = p.lr * p.regA * s1 * p.eps;
                                const float alpha = 1.f - 2.f * bb * inp;
                                ca[a] = aa; calpha[a] = alpha; cbd[a] = bb * ind;
                                np2 = cu1 * cu1 * (alpha * alpha * np2 + 2.f * alpha * aa * w[a] + aa * aa * G[a][a]);
#pragma unroll
                                for (int c = a + 1; c < kBlkK; ++c) w[c] = cu1 * (alpha * w[c] + aa * G[a][c]);
                            }
                            loss += (double)lsum;
#pragma unroll
                            for (int a = 0; a < kBlkK; ++a) {
                                if (a < nbk) {
                                    float di[V], dj[V];
#pragma unroll
                                    for (int v = 0; v < V; ++v) {
                                        const float step = fmaf(ca[a], pu[v], -cbd[a] * d[a][v]);      // a P - (b/|d|) d, old P
                                        const float pn = fmaf(calpha[a], pu[v], ca[a] * d[a][v]);
                                        di[v] = fmaf(-p.c_i, cur.qi[a][v] + step, step);
                                        dj[v] = fmaf(-p.c_i, cur.qj[a][v] - step, -step);
                                        pu[v] = fmaf(-p.c_u, pn, pn);
                                    }
                                    redq(reinterpret_cast<float*>(reinterpret_cast<char*>(cur.pi[a]) + ((warp & 1) ? cur.dx[a] : 0)), di);
                                    redq(cur.pj[a], dj);
                                }
                            }
                        }
                    }
                };
                if constexpr (V <= 2) {
                    if (flip) half(rowsB, rowsA); else half(rowsA, rowsB);
                    flip = !flip;
                } else {                                     // 128 floats per row: two bodies do not fit the register file (spills)
                    half(rowsA, rowsB);
                    rowsA = rowsB;
                }
            }
            if (!has_next) break;
            // ---- switch to segment seg+1: its first block is already in flight ------------------
            if (n_u != cur_u) {
                flush_user();
                cur_u = n_u;
#pragma unroll
                for (int v = 0; v < V; ++v) pu[v] = pu0[v] = pun[v];
            } else if (n_resync) {
                sync_user(n_u);       // a user shared between warps is re-read at every segment
            }
            u = n_u; len = n_len; resync = n_resync; my_i = n_i; my_j = n_j;
        }
        item = next_item;
    }
    flush_user();
    if (lane == 0 && loss != 0.0) atomicAdd(p.loss, loss);
}

}  // namespace yue
