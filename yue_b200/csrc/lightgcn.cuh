// K9: LightGCN training steps -- replaces the TF-1 graph of recommender/advanced/LightGCN.py:27-98 on
// base/DeepRecommender:22-35: per batch of `batch_size` consecutive training events, THREE sparse products over the whole
// play graph forward, the BPR loss on the propagated rows, the same products backward, and a dense Adam step on U and V.
//
//   e_0 = [U; V]      e_k = A e_{k-1}      F = e_0 + sum_k e_k * rsqrt(max(|e_k|^2, 1e-12))        (LightGCN.py:37-45)
//   A[u, m+t] = A[m+t, u] = c_ut^2: the SparseTensor holds one entry of value c_ut PER EVENT, duplicates add up (29-33)
//   loss = -sum log sigmoid(F_u . (F_i - F_j)) + regU/2 (|F_u|^2 + |F_i|^2 + |F_j|^2)   over the batch           (83-87)
//   Adam(lr): beta 0.9 / 0.999, epsilon 1e-8 outside the root, every row of U and V every step                    (88-90)
//
// B200 shape.  The reference's working sets are small (its own config: a 100K-record log): a step is eight short
// graph-wide phases, so what bounds it is launch and synchronisation latency, not bandwidth.  The whole range of steps is
// therefore ONE persistent cooperative launch (CTAs of 512 threads, as many as are resident at once), the phases separated by grid
// barriers (2 L + 2 per step), nothing returns to the host between steps, the negative sampler (K1's Philox stream, slot 4:
// the reference keeps the fifth of five draws, 68-75) is fused.  Per phase:
//   forward k = 1..L   e_k = A e_{k-1}, and the row's 1/norm while the row is in registers
//   batch              a lane group per triplet: F rows from the layers, loss, the three row gradients -> slot rows
//   combine            a lane group per slot: the FIRST slot of a graph row sums the row's slots in slot order (no atomics:
//                      same bits on every run), applies the norm-backward of every layer (NB_k) and stamps the row
//   backward k = L-1..0   D_k = A D_{k+1} + NB_k on stamped rows; D_L = NB_L exists only on the <= 3 B stamped rows, so the
//                      first product skips every other neighbour; k = 0 adds the row's own gradient and does the Adam
//                      update of the row in the same pass (D_0 is never stored)
// A row of the graph is worked on by a group of G lanes (G = 8 / 16 / 32 for rows of up to 32 / 64 / 128 floats, 16-byte
// chunks per lane, two chunks per lane up to 256), neighbour ids are read G at a time and broadcast, eight neighbour rows
// are in flight per group, the groups of a warp walk consecutive items in lockstep.  Rows with more than kGcnHeavy
// neighbours (the most played tracks, the heaviest users) are cut into chunks dealt to groups all over the grid; the group
// that arrives last adds the row's partial sums up in chunk order.  Light rows come in segments of consecutive rows whose
// neighbour lists are one range of the CSR (gcn_product, gcn_walk).  What was measured on the way: profiles/ncu_gcn_r2.md.
// Everything a later phase reads was written in an earlier one by other SMs: those loads are ld.global.cg (L2, never a
// stale L1 line); only the graph structure and the events go through the read-only path.
#pragma once
#include <cooperative_groups.h>
#include <cstdint>
#include <cuda_runtime.h>

#include "philox.cuh"

namespace yue {
namespace cg = cooperative_groups;

constexpr int kGcnMaxLayers = 4;
constexpr int kGcnThreads = 512;
constexpr int kGcnHeavy = 64;          // neighbours beyond which a CTA shares the row
constexpr int kGcnSegRows = 8;         // light rows per segment (<= the smallest group), at most kGcnSegEdges neighbours together
constexpr int kGcnSegEdges = 64;
constexpr int kGcnHashSize = 8192;     // the stamped rows' hash table: >= 2 x kGcnSmemSlots, a power of two
constexpr int kGcnSmemSlots = 3072;    // slot ids kept in shared memory by the combine phase (batches of up to 1024)
constexpr int kGcnMaxBatch = 4096;
constexpr uint32_t kGcnNegSlot = 4;    // the fifth draw is the one LightGCN.py:76-78 keeps
constexpr float kGcnNormEps = 1e-12f;

// a chunk of a heavy row's neighbours: [off, off + len) of row `row`; its partial sum goes to partial[slot]; the row's chunks
// are partial[first_slot .. first_slot + n_chunks) and count themselves in at arrived[hidx]
struct GcnChunk { int32_t row, off, len, slot, first_slot, n_chunks, hidx, pad; };

struct GcnParams {
    int64_t m, n;                      // users, tracks; graph rows: users first, then tracks (M = m + n)
    int ld, L;
    const int64_t* u_indptr; const int32_t* u_items; const int32_t* u_cnt;   // user rows: tracks and play counts
    const int64_t* t_indptr; const int32_t* t_users; const int32_t* t_cnt;   // track rows: users and play counts
    float* P; float* Q;                // e_0 = the variables U, V
    float* E[kGcnMaxLayers];           // e_1 .. e_L, [M, ld]
    float* rinv;                       // [L][M] 1 / max(norm, 1e-6) of e_k's rows
    float* D[2];                       // backward ping-pong, [M, ld]
    float* am; float* av;              // Adam moments, [M, ld]
    uint32_t* stamp; int32_t* slot_of; // [M]: the step that last touched the row, and its leading slot
    int32_t* slot_row; float* slot_grad;   // [3 B], [3 B, ld]
    float* NB;                         // [L + 1][3 B, ld]: NB_0 = the row's summed gradient, NB_k = norm-backward of layer k
    float* trip_loss;                  // [B]
    double* loss_out;                  // [steps of this launch]
    const int32_t* ev_user; const int32_t* ev_item; const int32_t* ev_neg;   // file order; ev_neg == nullptr: draw them
    int64_t T; int batch; int64_t step_begin, step_end;
    uint64_t seed; uint32_t epoch; uint32_t stamp_base;
    int64_t adam_t;                    // Adam steps taken before this launch
    float lr, reg;
    const GcnChunk* chunks; int64_t n_chunk;   // heavy rows cut into chunks of neighbours
    float* partial; unsigned* arrived;         // [n_chunk, ld] partial sums, [heavy rows] chunks that have arrived
    const int2* segs; int64_t n_seg;           // light rows: (first row, rows) of up to kGcnSegRows consecutive rows of one side
    float* FP; float* FQ;              // forward_only: where F goes
    unsigned long long* phase_ns;      // diagnostics (YUE_GCN_TIMING): [16] nanoseconds per phase, summed over the steps, by CTA 0
    int forward_only;
};

template <int G> __device__ __forceinline__ unsigned gcn_group_mask() {
    if (G == 32) return 0xffffffffu;
    return ((1u << (G & 31)) - 1u) << ((threadIdx.x & 31) & ~(G - 1));
}
template <int G> __device__ __forceinline__ float gcn_group_sum(float x, unsigned mask) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) x += __shfl_xor_sync(mask, x, o);
    return x;
}
template <int G> __device__ __forceinline__ double gcn_group_sum(double x, unsigned mask) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) x += __shfl_xor_sync(mask, x, o);
    return x;
}

template <int G, int NC> struct GcnVec {
    float4 c[NC];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int q = 0; q < NC; ++q) c[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // rows written earlier in this launch: L2 loads
    __device__ __forceinline__ void load(const float* row, int ld, int gl) {
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            const int f = 4 * (gl + G * q);
            c[q] = f < ld ? __ldcg(reinterpret_cast<const float4*>(row + f)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __device__ __forceinline__ void store(float* row, int ld, int gl) const {
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            const int f = 4 * (gl + G * q);
            if (f < ld) *reinterpret_cast<float4*>(row + f) = c[q];
        }
    }
    __device__ __forceinline__ void axpy(float a, const GcnVec& x) {
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            c[q].x = fmaf(a, x.c[q].x, c[q].x); c[q].y = fmaf(a, x.c[q].y, c[q].y);
            c[q].z = fmaf(a, x.c[q].z, c[q].z); c[q].w = fmaf(a, x.c[q].w, c[q].w);
        }
    }
    __device__ __forceinline__ void add(const GcnVec& x) { axpy(1.f, x); }
    __device__ __forceinline__ float dot_part(const GcnVec& x) const {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < NC; ++q) s += c[q].x * x.c[q].x + c[q].y * x.c[q].y + c[q].z * x.c[q].z + c[q].w * x.c[q].w;
        return s;
    }
};

// where a graph row's neighbours are, and on which side of the tables they live
struct GcnRowRange { const int32_t* nbr; const int32_t* cnt; int64_t b, e; int side; };   // side 0: a user row (neighbours are tracks)
__device__ __forceinline__ GcnRowRange gcn_row_range(const GcnParams& p, int64_t r) {
    GcnRowRange x;
    if (r < p.m) { x.nbr = p.u_items; x.cnt = p.u_cnt; x.b = __ldg(p.u_indptr + r); x.e = __ldg(p.u_indptr + r + 1); x.side = 0; }
    else { const int64_t t = r - p.m; x.nbr = p.t_users; x.cnt = p.t_cnt; x.b = __ldg(p.t_indptr + t); x.e = __ldg(p.t_indptr + t + 1); x.side = 1; }
    return x;
}

// a table of graph rows given as its user part and its track part (e_0 is two allocations, every other table one)
struct GcnTable {
    float* users; float* tracks; int ld; int64_t m;
    __device__ __forceinline__ float* row(int64_t r) const { return r < m ? users + r * ld : tracks + (r - m) * ld; }
    __device__ __forceinline__ const float* nbr_row(int side, int32_t id) const { return (side == 0 ? tracks : users) + (int64_t)id * ld; }
};
__device__ __forceinline__ GcnTable gcn_table(const GcnParams& p, float* x) { return GcnTable{x, x + p.m * p.ld, p.ld, p.m}; }
__device__ __forceinline__ GcnTable gcn_table0(const GcnParams& p) { return GcnTable{p.P, p.Q, p.ld, p.m}; }

// sources of a product: a dense table, or NB_L on the stamped rows only
template <int G, int NC> struct GcnDenseSrc {
    GcnTable t;
    __device__ __forceinline__ void fetch(int side, int32_t id, int gl, GcnVec<G, NC>& v) const { v.load(t.nbr_row(side, id), t.ld, gl); }
};
// NB_L on the rows the batch stamped.  The stamped rows (<= 3 B of up to a million) sit in a hash table in shared memory
// that every CTA builds from the slot list at the start of the phase (row -> leading slot + 1, the row itself is slot_ids[slot]): a neighbour that is not in it
// costs a shared-memory probe instead of a load from L2.  Batches beyond kGcnSmemSlots slots use the stamps in global memory.
__device__ __forceinline__ uint32_t gcn_hash(int32_t r) { return ((uint32_t)r * 2654435761u) >> 19; }   // 13 bits
template <int G, int NC> struct GcnStampedSrc {
    const uint32_t* stamp; const int32_t* slot_of; const float* nb; uint32_t cur; int ld; int64_t m;
    const int* table; const int32_t* slot_ids;                      // shared memory; table == nullptr: use the stamps
    __device__ __forceinline__ void fetch(int side, int32_t id, int gl, GcnVec<G, NC>& v) const {
        const int64_t r = side == 0 ? m + id : id;
        if (table) {
            uint32_t hsh = gcn_hash((int32_t)r);
            int e = table[hsh];
            while (e != 0 && slot_ids[e - 1] != (int32_t)r) { hsh = (hsh + 1) & (kGcnHashSize - 1); e = table[hsh]; }
            if (e != 0) v.load(nb + (int64_t)(e - 1) * ld, ld, gl); else v.zero();
            return;
        }
        if (__ldcg(stamp + r) == cur) v.load(nb + (int64_t)__ldcg(slot_of + r) * ld, ld, gl);
        else v.zero();
    }
};

// neighbour rows in flight per group (registers: 4 NC floats each)
template <int NC> struct GcnFlight { static constexpr int value = NC == 1 ? 8 : 4; };

// Walk the neighbours [eb, eb + n_edges) of `nr` consecutive rows (lane x holds where row x ends, relative to eb) with
// GcnFlight rows in flight ACROSS row boundaries; the ids of the next G neighbours are fetched while the current ones are
// worked on.  fin.begin(row) when a row starts, fin.finish(row, acc) when it is complete (rows without neighbours too).
template <int G, int NC, class Src, class Fin>
__device__ __forceinline__ void gcn_walk(const GcnRowRange& rr, int64_t eb, int n_edges, int nr, int my_end, int gl, unsigned mask,
                                         const Src& src, Fin& fin) {
    // The groups of a warp walk their items IN LOCKSTEP: the trip counts below are the warp's maximum, a group that is
    // through its own neighbours runs empty iterations.  Two groups that drift apart are issued one after the other (the
    // warp's instruction stream doubles) and every 16-lane shuffle becomes a MATCH / VOTE sequence (profiles/ncu_gcn_r2.md);
    // in lockstep the shuffles of the walk are plain full-warp SHFLs with a width.  Only the end of a row (finish_row) is
    // group-uniform code: its shuffles name the group's lanes, and the warp reconverges behind it.
    constexpr int FL = GcnFlight<NC>::value;
    constexpr unsigned full = 0xffffffffu;
    int n_iter = (n_edges + G - 1) / G;
#pragma unroll
    for (int o = G; o < 32; o <<= 1) n_iter = max(n_iter, __shfl_xor_sync(full, n_iter, o));
    const int first_end = __shfl_sync(full, my_end, 0, G);            // every lane of the warp executes it: an empty item too
    int row = 0, row_end = nr > 0 ? first_end : 0x7fffffff;
    GcnVec<G, NC> acc; acc.zero();
    if (nr > 0) fin.begin(0);
    auto finish_row = [&]() {
        fin.finish(row, acc);
        acc.zero(); ++row;
        if (row < nr) { row_end = __shfl_sync(mask, my_end, row, G); fin.begin(row); }
        else row_end = 0x7fffffff;
    };
    int32_t id_n = 0, cn_n = 0;
    if (gl < n_edges) { id_n = __ldg(rr.nbr + eb + gl); cn_n = __ldg(rr.cnt + eb + gl); }
    for (int c = 0; c < n_iter; ++c) {
        const int p0 = c * G;
        const int32_t id = id_n; const float w = (float)cn_n * (float)cn_n;
        id_n = 0; cn_n = 0;
        if (p0 + G + gl < n_edges) { id_n = __ldg(rr.nbr + eb + p0 + G + gl); cn_n = __ldg(rr.cnt + eb + p0 + G + gl); }
        const int cnt = n_edges - p0 < G ? n_edges - p0 : G;           // <= 0: this group is through
#pragma unroll 1
        for (int t = 0; t < G; t += FL) {
            GcnVec<G, NC> v[FL]; float ww[FL];
#pragma unroll
            for (int x = 0; x < FL; ++x) {
                const int32_t nid = __shfl_sync(full, id, t + x, G);
                ww[x] = __shfl_sync(full, w, t + x, G);
                if ((t + x) < cnt) src.fetch(rr.side, nid, gl, v[x]); else v[x].zero();
            }
#pragma unroll
            for (int x = 0; x < FL; ++x) {
                if ((t + x) < cnt) {
                    while (p0 + t + x >= row_end) finish_row();
                    acc.axpy(ww[x], v[x]);
                }
            }
        }
    }
    while (row < nr) finish_row();
}

// One product phase: epi(row, acc) for every row of the graph.  Work items, dealt to the groups of the whole grid in turn:
//  * CHUNKS of heavy rows (more than kGcnHeavy neighbours; a chunk is ~sqrt(degree), at least kGcnHeavy, neighbours): the
//    group leaves its partial sum in scratch and counts itself in; the group that arrives LAST adds the row's partials up
//    in chunk order (same bits on every run, whoever is last) and runs the row's epilogue.  No CTA waits for another.
//  * SEGMENTS of up to kGcnSegRows consecutive light rows: their neighbour lists are one contiguous range of the CSR; a light
//    row has ~2 neighbours, row by row each would cost its own chain of dependent latencies.
template <int G, int NC, class Src, class Epi>
__device__ __forceinline__ void gcn_product(const GcnParams& p, const Src& src, const Epi& epi, uint32_t cur, int* cta_next) {
    constexpr int NGRP = kGcnThreads / G, FL = GcnFlight<NC>::value;
    const int gl = threadIdx.x % G;
    const unsigned mask = gcn_group_mask<G>();
    // Items (largest first) are dealt to the CTAs in runs of NGRP consecutive items -- neighbouring groups work on neighbouring
    // rows, and every CTA gets about the same work -- and inside a CTA a WARP that is done takes the CTA's next 32 / G items, one
    // per group (a counter in shared memory; a single global ticket counter was tried: 12 K same-address atomics per phase cost
    // more than the imbalance they removed).  The chunk list is padded to a multiple of 32 / G, so the groups of a warp always
    // work on items of one kind, side by side (gcn_walk).
    constexpr int GPW = 32 / G, NWARP = kGcnThreads / 32;
    const int sub = (threadIdx.x & 31) / G, warp = threadIdx.x >> 5;
    const int64_t n_items = p.n_chunk + p.n_seg, tot = (int64_t)gridDim.x * NGRP;
    if (threadIdx.x == 0) *cta_next = NWARP;
    __syncthreads();
    auto first_item = [&](int k) { return (int64_t)(k / NWARP) * tot + (int64_t)blockIdx.x * NGRP + (int64_t)(k % NWARP) * GPW; };
    auto take = [&]() -> int64_t {
        int t = 0;
        if ((threadIdx.x & 31) == 0) t = atomicAdd(cta_next, 1);
        return first_item(__shfl_sync(0xffffffffu, t, 0));
    };
    for (int64_t it0 = first_item(warp); it0 < n_items; it0 = take()) {
        const int64_t it = it0 + sub;
        if (it < p.n_chunk) {
            const GcnChunk ck = p.chunks[it];
            const GcnRowRange rr = gcn_row_range(p, ck.row);
            struct {
                const GcnParams& p; const Epi& epi; const GcnChunk& ck; uint32_t cur; int gl; unsigned mask;
                __device__ __forceinline__ void begin(int) {}
                __device__ __forceinline__ void finish(int, GcnVec<G, NC>& acc) {
                    const int ld = p.ld;
                    acc.store(p.partial + (int64_t)ck.slot * ld, ld, gl);
                    __threadfence();
                    __syncwarp(mask);
                    unsigned old = 0;
                    if (gl == 0) old = atomicAdd(p.arrived + ck.hidx, 1u);
                    old = __shfl_sync(mask, old, 0, G);
                    if (old != (unsigned)ck.n_chunks - 1u) return;
                    __threadfence();                                     // last to arrive: every partial of the row is visible
                    typename Epi::Pre pre; epi.prefetch(ck.row, gl, pre);
                    const bool stamped = Epi::kStamp && __ldcg(p.stamp + ck.row) == cur;
                    acc.zero();
                    for (int c0 = 0; c0 < ck.n_chunks; c0 += FL) {
                        GcnVec<G, NC> v[FL];
#pragma unroll
                        for (int x = 0; x < FL; ++x) { if (c0 + x < ck.n_chunks) v[x].load(p.partial + (int64_t)(ck.first_slot + c0 + x) * ld, ld, gl); else v[x].zero(); }
#pragma unroll
                        for (int x = 0; x < FL; ++x) acc.add(v[x]);
                    }
                    epi(ck.row, acc, pre, stamped, gl, mask);
                    if (gl == 0) p.arrived[ck.hidx] = 0u;                // the next phase counts from zero
                }
            } fin{p, epi, ck, cur, gl, mask};
            gcn_walk<G, NC>(rr, rr.b + ck.off, ck.len, 1, ck.len, gl, mask, src, fin);
        } else {
            // first row, rows (consecutive, one side); the last run of a phase may be short: an empty item, through the same code
            const int2 sg = it < n_items ? __ldg(p.segs + (it - p.n_chunk)) : make_int2(0, 0);
            const int64_t r0 = sg.x; const int nr = sg.y;
            GcnRowRange rr;
            const int64_t* ip;
            if (r0 < p.m) { rr.nbr = p.u_items; rr.cnt = p.u_cnt; rr.side = 0; ip = p.u_indptr + r0; }
            else { rr.nbr = p.t_users; rr.cnt = p.t_cnt; rr.side = 1; ip = p.t_indptr + (r0 - p.m); }
            const int64_t eb = __ldg(ip);
            const int my_end = gl < nr ? (int)(__ldg(ip + 1 + gl) - eb) : 0; // lane x: where row r0 + x ends, from eb
            const uint32_t my_stamp = Epi::kStamp && gl < nr ? __ldcg(p.stamp + r0 + gl) : 0u;
            const int n_edges = __shfl_sync(mask, my_end, nr - 1, G);
            struct {
                const Epi& epi; int64_t r0; uint32_t my_stamp, cur; int gl; unsigned mask; typename Epi::Pre pre;
                __device__ __forceinline__ void begin(int row) { epi.prefetch(r0 + row, gl, pre); }
                __device__ __forceinline__ void finish(int row, GcnVec<G, NC>& acc) {
                    const bool stamped = Epi::kStamp && __shfl_sync(mask, my_stamp, row, G) == cur;
                    epi(r0 + row, acc, pre, stamped, gl, mask);
                }
            } fin{epi, r0, my_stamp, cur, gl, mask, {}};
            gcn_walk<G, NC>(rr, eb, n_edges, nr, my_end, gl, mask, src, fin);
        }
    }
}

template <int G, int NC> struct GcnForwardEpi {
    static constexpr bool kStamp = false;
    struct Pre {};
    float* out; float* rinv; int ld;
    __device__ __forceinline__ void prefetch(int64_t, int, Pre&) const {}
    __device__ __forceinline__ void operator()(int64_t r, GcnVec<G, NC>& acc, const Pre&, bool, int gl, unsigned mask) const {
        acc.store(out + r * ld, ld, gl);
        const float ss = gcn_group_sum<G>(acc.dot_part(acc), mask);
        if (gl == 0) rinv[r] = rsqrtf(fmaxf(ss, kGcnNormEps));
    }
};

// D_k = (product) + NB_k on stamped rows, k >= 1
template <int G, int NC> struct GcnBackwardEpi {
    static constexpr bool kStamp = true;
    struct Pre {};
    const int32_t* slot_of; const float* nb_k; int ld; float* out;
    __device__ __forceinline__ void prefetch(int64_t, int, Pre&) const {}
    __device__ __forceinline__ void operator()(int64_t r, GcnVec<G, NC>& acc, const Pre&, bool stamped, int gl, unsigned) const {
        if (stamped) {
            GcnVec<G, NC> x; x.load(nb_k + (int64_t)__ldcg(slot_of + r) * ld, ld, gl);
            acc.add(x);
        }
        acc.store(out + r * ld, ld, gl);
    }
};

// k = 0: the gradient of the row of e_0 is complete in registers -- Adam's step on the row, D_0 is never stored
template <int G, int NC> struct GcnAdamEpi {
    static constexpr bool kStamp = true;
    struct Pre { GcnVec<G, NC> mo, ve, va; };
    const int32_t* slot_of; const float* nb_0; int ld;
    GcnTable var; float* am; float* av; float lr_t;
    __device__ __forceinline__ void prefetch(int64_t r, int gl, Pre& pre) const {
        pre.mo.load(am + r * ld, ld, gl); pre.ve.load(av + r * ld, ld, gl); pre.va.load(var.row(r), ld, gl);
    }
    __device__ __forceinline__ void operator()(int64_t r, GcnVec<G, NC>& acc, Pre& pre, bool stamped, int gl, unsigned) const {
        if (stamped) {
            GcnVec<G, NC> x; x.load(nb_0 + (int64_t)__ldcg(slot_of + r) * ld, ld, gl);
            acc.add(x);
        }
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            float* g = &acc.c[q].x; float* mm = &pre.mo.c[q].x; float* vv = &pre.ve.c[q].x; float* xx = &pre.va.c[q].x;
#pragma unroll
            for (int z = 0; z < 4; ++z) {
                mm[z] = 0.9f * mm[z] + 0.1f * g[z];
                vv[z] = 0.999f * vv[z] + 0.001f * g[z] * g[z];
                xx[z] -= lr_t * mm[z] / (sqrtf(vv[z]) + 1e-8f);
            }
        }
        pre.mo.store(am + r * ld, ld, gl); pre.ve.store(av + r * ld, ld, gl); pre.va.store(var.row(r), ld, gl);
    }
};

// F row of graph row r: e_0 + sum_k e_k / max(norm, 1e-6)
template <int G, int NC>
__device__ __forceinline__ void gcn_final_row(const GcnParams& p, int64_t r, int gl, GcnVec<G, NC>& f) {
    const int64_t M = p.m + p.n;
    f.load(gcn_table0(p).row(r), p.ld, gl);
    for (int k = 1; k <= p.L; ++k) {
        GcnVec<G, NC> e; e.load(p.E[k - 1] + r * p.ld, p.ld, gl);
        f.axpy(__ldcg(p.rinv + (int64_t)(k - 1) * M + r), e);
    }
}

__device__ __forceinline__ unsigned long long gcn_now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// phase slots of GcnParams::phase_ns: 0..3 forward k, 4 batch, 5 combine, 6..9 backward k, then the same + 16 for the barrier after it
#define GCN_TICK(slot)                                                                          \
    if (p.phase_ns && blockIdx.x == 0 && threadIdx.x == 0) { const unsigned long long t1_ = gcn_now(); p.phase_ns[slot] += t1_ - t_last; t_last = t1_; }

template <int G, int NC>
__global__ void __launch_bounds__(kGcnThreads, 1) gcn_steps_kernel(const GcnParams p) {
    unsigned long long t_last = p.phase_ns ? gcn_now() : 0ull;
    constexpr int NGRP = kGcnThreads / G;
    __shared__ int32_t slot_ids[kGcnSmemSlots];        // the batch's slot -> graph row
    __shared__ int cta_next;                           // product phases: the CTA's next unclaimed item
    __shared__ int hash_tab[kGcnHashSize];             // stamped row -> its leading slot + 1 (0: empty), open addressing
    cg::grid_group grid = cg::this_grid();
    const int gl = threadIdx.x % G, grp = threadIdx.x / G;
    const unsigned mask = gcn_group_mask<G>();
    const int64_t M = p.m + p.n, ggrp = (int64_t)blockIdx.x * NGRP + grp, tot = (int64_t)gridDim.x * NGRP;
    const int ld = p.ld;
    const int64_t steps = p.forward_only ? 1 : p.step_end - p.step_begin;

    for (int64_t s = 0; s < steps; ++s) {
        // ---- forward: e_k = A e_{k-1} ----
        for (int k = 1; k <= p.L; ++k) {
            const GcnDenseSrc<G, NC> src{k == 1 ? gcn_table0(p) : gcn_table(p, p.E[k - 2])};
            const GcnForwardEpi<G, NC> epi{p.E[k - 1], p.rinv + (int64_t)(k - 1) * M, ld};
            gcn_product<G, NC>(p, src, epi, 0u, &cta_next);
            GCN_TICK(k - 1)
            grid.sync();
            GCN_TICK(16 + k - 1)
        }
        if (p.forward_only) {
            for (int64_t r = ggrp; r < M; r += tot) {
                GcnVec<G, NC> f; gcn_final_row<G, NC>(p, r, gl, f);
                f.store(r < p.m ? p.FP + r * ld : p.FQ + (r - p.m) * ld, ld, gl);
            }
            return;
        }
        const int64_t step = p.step_begin + s, ev0 = step * p.batch;
        const int nb = (int)((p.T - ev0) < (int64_t)p.batch ? (p.T - ev0) : (int64_t)p.batch);
        const uint32_t cur = p.stamp_base + (uint32_t)s + 1u;

        // ---- batch: loss and the three row gradients of every triplet ----
        for (int64_t b = ggrp; b < nb; b += tot) {
            const int64_t e = ev0 + b;
            const int32_t u = __ldg(p.ev_user + e), i = __ldg(p.ev_item + e);
            int32_t j;
            if (p.ev_neg) j = __ldg(p.ev_neg + e);
            else {
                const int64_t rb = __ldg(p.u_indptr + u);
                j = sample_negative(p.seed, p.epoch, (uint64_t)e, kGcnNegSlot, (uint32_t)p.n, p.u_items + rb, (int)(__ldg(p.u_indptr + u + 1) - rb));
            }
            GcnVec<G, NC> fu, fi, fj;
            gcn_final_row<G, NC>(p, u, gl, fu);
            gcn_final_row<G, NC>(p, p.m + i, gl, fi);
            gcn_final_row<G, NC>(p, p.m + j, gl, fj);
            const float y = gcn_group_sum<G>(fu.dot_part(fi) - fu.dot_part(fj), mask);
            const float n2 = gcn_group_sum<G>(fu.dot_part(fu) + fi.dot_part(fi) + fj.dot_part(fj), mask);
            const float c = 1.f / (1.f + expf(y));                                  // 1 - sigmoid(y)
            const float nll = y > 0.f ? log1pf(expf(-y)) : -y + log1pf(expf(y));    // -log sigmoid(y)
            GcnVec<G, NC> gu, gi, gj;
            gu.zero(); gu.axpy(p.reg, fu); gu.axpy(-c, fi); gu.axpy(c, fj);
            gi.zero(); gi.axpy(p.reg, fi); gi.axpy(-c, fu);
            gj.zero(); gj.axpy(p.reg, fj); gj.axpy(c, fu);
            gu.store(p.slot_grad + (3 * b + 0) * ld, ld, gl);
            gi.store(p.slot_grad + (3 * b + 1) * ld, ld, gl);
            gj.store(p.slot_grad + (3 * b + 2) * ld, ld, gl);
            if (gl == 0) {
                p.slot_row[3 * b + 0] = u; p.slot_row[3 * b + 1] = (int32_t)(p.m + i); p.slot_row[3 * b + 2] = (int32_t)(p.m + j);
                p.trip_loss[b] = nll + 0.5f * p.reg * n2;
            }
        }
        GCN_TICK(4)
        grid.sync();
        GCN_TICK(20)

        // ---- combine: the first slot of a row sums the row's slots in slot order, norm-backward per layer, stamp ----
        const int ns = 3 * nb;
        const bool ids_in_smem = ns <= kGcnSmemSlots;                 // the slots' rows: every leader test scans them
        if (ids_in_smem) {
            for (int x = threadIdx.x; x < ns; x += kGcnThreads) slot_ids[x] = __ldcg(p.slot_row + x);
            __syncthreads();
        }
        auto slot_id = [&](int64_t x) { return ids_in_smem ? slot_ids[x] : __ldcg(p.slot_row + x); };
        for (int64_t sl = ggrp; sl < ns; sl += tot) {
            const int32_t r = slot_id(sl);
            bool dup = false;
            for (int64_t x = gl; x < sl; x += G) dup |= slot_id(x) == r;
            if (!__any_sync(mask, dup)) {
                GcnVec<G, NC> g; g.load(p.slot_grad + sl * ld, ld, gl);
                for (int64_t x0 = sl + 1; x0 < ns; x0 += G) {
                    const bool hit = x0 + gl < ns && slot_id(x0 + gl) == r;
                    unsigned bal = (__ballot_sync(mask, hit) & mask) >> ((threadIdx.x & 31) & ~(G - 1));
                    while (bal) {
                        const int t = __ffs(bal) - 1; bal &= bal - 1;
                        GcnVec<G, NC> x; x.load(p.slot_grad + (x0 + t) * ld, ld, gl);
                        g.add(x);
                    }
                }
                g.store(p.NB + sl * ld, ld, gl);
                for (int k = 1; k <= p.L; ++k) {
                    GcnVec<G, NC> e; e.load(p.E[k - 1] + (int64_t)r * ld, ld, gl);
                    const float ri = __ldcg(p.rinv + (int64_t)(k - 1) * M + r);
                    const float ss = gcn_group_sum<G>(e.dot_part(e), mask);
                    GcnVec<G, NC> nbk; nbk.zero(); nbk.axpy(ri, g);
                    if (ss > kGcnNormEps) {                                      // d/de of e * rsqrt(|e|^2): ri (g - n (n . g)), n = e ri
                        const float dt = gcn_group_sum<G>(e.dot_part(g), mask) * ri;      // n . g
                        nbk.axpy(-dt * ri * ri, e);
                    }
                    nbk.store(p.NB + ((int64_t)k * 3 * p.batch + sl) * ld, ld, gl);
                }
                if (gl == 0) { p.stamp[r] = cur; p.slot_of[r] = (int32_t)sl; }
            }
            if (sl == 0) {                                                        // the batch's loss, summed in a fixed order
                double a = 0.0;
                for (int x = gl; x < nb; x += G) a += (double)__ldcg(p.trip_loss + x);
                a = gcn_group_sum<G>(a, mask);
                if (gl == 0) p.loss_out[s] = a;
            }
        }
        GCN_TICK(5)
        grid.sync();
        GCN_TICK(21)

        // ---- backward: D_k = A D_{k+1} + NB_k, k = L-1 .. 0; k = 0 is the Adam step ----
        const double t_adam = (double)(p.adam_t + s + 1);
        const float lr_t = (float)((double)p.lr * sqrt(1.0 - pow(0.999, t_adam)) / (1.0 - pow(0.9, t_adam)));
        for (int k = p.L - 1; k >= 0; --k) {
            const unsigned long long t_adam0 = p.phase_ns && k == 0 ? gcn_now() : 0ull;
            const float* nb_k = p.NB + (int64_t)k * 3 * p.batch * ld;
            const GcnBackwardEpi<G, NC> epi{p.slot_of, nb_k, ld, p.D[k & 1]};
            const GcnAdamEpi<G, NC> adam{p.slot_of, nb_k, ld, gcn_table0(p), p.am, p.av, lr_t};
            if (k == p.L - 1) {
                if (ids_in_smem) {                                   // stamped row -> leading slot, built by every CTA for itself
                    for (int x = threadIdx.x; x < kGcnHashSize; x += kGcnThreads) hash_tab[x] = 0;
                    __syncthreads();
                    for (int x = threadIdx.x; x < ns; x += kGcnThreads) {
                        const int32_t r = slot_ids[x];
                        if (__ldcg(p.slot_of + r) == x) {
                            uint32_t hsh = gcn_hash(r);
                            while (atomicCAS(&hash_tab[hsh], 0, x + 1) != 0) hsh = (hsh + 1) & (kGcnHashSize - 1);
                        }
                    }
                    __syncthreads();
                }
                const GcnStampedSrc<G, NC> src{p.stamp, p.slot_of, p.NB + (int64_t)p.L * 3 * p.batch * ld, cur, ld, p.m,
                                               ids_in_smem ? hash_tab : nullptr, slot_ids};
                if (k > 0) gcn_product<G, NC>(p, src, epi, cur, &cta_next); else gcn_product<G, NC>(p, src, adam, cur, &cta_next);
            } else {
                const GcnDenseSrc<G, NC> src{gcn_table(p, p.D[(k + 1) & 1])};
                if (k > 0) gcn_product<G, NC>(p, src, epi, cur, &cta_next); else gcn_product<G, NC>(p, src, adam, cur, &cta_next);
            }
            GCN_TICK(6 + k)
            if (p.phase_ns && k == 0 && threadIdx.x == 0) p.phase_ns[32 + blockIdx.x] += gcn_now() - t_adam0;   // per CTA: its Adam phase
            grid.sync();
            GCN_TICK(22 + k)
        }
    }
}

}  // namespace yue
