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
// are in flight per group.  Rows with more than kGcnHeavy neighbours (the most played tracks, the heaviest users) are
// worked on by a whole CTA: every group takes a slice, partial sums meet in shared memory in a fixed order.
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
constexpr int kGcnMaxBatch = 4096;
constexpr uint32_t kGcnNegSlot = 4;    // the fifth draw is the one LightGCN.py:76-78 keeps
constexpr float kGcnNormEps = 1e-12f;

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
    const int32_t* heavy_rows; int n_heavy; const int32_t* light_rows; int64_t n_light;
    float* FP; float* FQ;              // forward_only: where F goes
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
template <int G, int NC> struct GcnStampedSrc {
    const uint32_t* stamp; const int32_t* slot_of; const float* nb; uint32_t cur; int ld; int64_t m;
    __device__ __forceinline__ void fetch(int side, int32_t id, int gl, GcnVec<G, NC>& v) const {
        const int64_t r = side == 0 ? m + id : id;
        if (__ldcg(stamp + r) == cur) v.load(nb + (int64_t)__ldcg(slot_of + r) * ld, ld, gl);
        else v.zero();
    }
};

// neighbour rows in flight per group (registers: 4 NC floats each)
template <int NC> struct GcnFlight { static constexpr int value = NC == 1 ? 8 : 4; };

// acc += sum over neighbours [b, e) of w * src(neighbour),  w = count^2
template <int G, int NC, class Src>
__device__ __forceinline__ void gcn_gather(const GcnRowRange& rr, int64_t b, int64_t e, int gl, unsigned mask, const Src& src, GcnVec<G, NC>& acc) {
    constexpr int kGcnFlight = GcnFlight<NC>::value;
    for (int64_t p0 = b; p0 < e; p0 += G) {
        int32_t id = 0; float w = 0.f;
        if (p0 + gl < e) { id = __ldg(rr.nbr + p0 + gl); const float c = (float)__ldg(rr.cnt + p0 + gl); w = c * c; }
        const int cnt = (int)((e - p0) < (int64_t)G ? (e - p0) : (int64_t)G);
        for (int t = 0; t < cnt; t += kGcnFlight) {
            GcnVec<G, NC> v[kGcnFlight]; float ww[kGcnFlight];
#pragma unroll
            for (int x = 0; x < kGcnFlight; ++x) {
                const int lane = (t + x) & (G - 1);                       // past the end: a lane whose weight is 0 (or id 0)
                const int32_t nid = __shfl_sync(mask, id, lane, G);
                const float wv = __shfl_sync(mask, w, lane, G);
                ww[x] = (t + x) < cnt ? wv : 0.f;
                if ((t + x) < cnt) src.fetch(rr.side, nid, gl, v[x]); else v[x].zero();
            }
#pragma unroll
            for (int x = 0; x < kGcnFlight; ++x) acc.axpy(ww[x], v[x]);
        }
    }
}

// one product phase: epi(row, acc) for every row of the graph
template <int G, int NC, class Src, class Epi>
__device__ __forceinline__ void gcn_product(const GcnParams& p, const Src& src, const Epi& epi, float* part) {
    constexpr int NGRP = kGcnThreads / G, W = 4 * G * NC;
    const int gl = threadIdx.x % G, grp = threadIdx.x / G;
    const unsigned mask = gcn_group_mask<G>();
    for (int hi = blockIdx.x; hi < p.n_heavy; hi += gridDim.x) {
        const int64_t r = __ldg(p.heavy_rows + hi);
        const GcnRowRange rr = gcn_row_range(p, r);
        const int64_t per = ((rr.e - rr.b + NGRP - 1) / NGRP + 7) / 8 * 8;
        const int64_t gb = rr.b + grp * per, ge = gb + per < rr.e ? gb + per : rr.e;
        GcnVec<G, NC> acc; acc.zero();
        if (gb < ge) gcn_gather<G, NC>(rr, gb, ge, gl, mask, src, acc);
#pragma unroll
        for (int q = 0; q < NC; ++q) *reinterpret_cast<float4*>(part + grp * W + 4 * (gl + G * q)) = acc.c[q];
        __syncthreads();
        if (grp == 0) {
            for (int g2 = 1; g2 < NGRP; ++g2) {
#pragma unroll
                for (int q = 0; q < NC; ++q) {
                    const float4 x = *reinterpret_cast<const float4*>(part + g2 * W + 4 * (gl + G * q));
                    acc.c[q].x += x.x; acc.c[q].y += x.y; acc.c[q].z += x.z; acc.c[q].w += x.w;
                }
            }
            epi(r, acc, gl, mask);
        }
        __syncthreads();
    }
    const int64_t tot = (int64_t)gridDim.x * NGRP;
    for (int64_t li = (int64_t)blockIdx.x * NGRP + grp; li < p.n_light; li += tot) {
        const int64_t r = __ldg(p.light_rows + li);
        const GcnRowRange rr = gcn_row_range(p, r);
        GcnVec<G, NC> acc; acc.zero();
        gcn_gather<G, NC>(rr, rr.b, rr.e, gl, mask, src, acc);
        epi(r, acc, gl, mask);
    }
}

template <int G, int NC> struct GcnForwardEpi {
    float* out; float* rinv; int ld;
    __device__ __forceinline__ void operator()(int64_t r, GcnVec<G, NC>& acc, int gl, unsigned mask) const {
        acc.store(out + r * ld, ld, gl);
        const float ss = gcn_group_sum<G>(acc.dot_part(acc), mask);
        if (gl == 0) rinv[r] = rsqrtf(fmaxf(ss, kGcnNormEps));
    }
};

template <int G, int NC> struct GcnBackwardEpi {
    const uint32_t* stamp; const int32_t* slot_of; const float* nb_k;    // NB_k (k = 0: the summed gradient itself)
    uint32_t cur; int ld;
    float* out;                                                           // D_k, or nullptr at k = 0
    GcnTable var; float* am; float* av; float lr_t;                       // k = 0: Adam on the row of e_0
    __device__ __forceinline__ void operator()(int64_t r, GcnVec<G, NC>& acc, int gl, unsigned) const {
        if (__ldcg(stamp + r) == cur) {
            GcnVec<G, NC> x; x.load(nb_k + (int64_t)__ldcg(slot_of + r) * ld, ld, gl);
            acc.add(x);
        }
        if (out) { acc.store(out + r * ld, ld, gl); return; }
        GcnVec<G, NC> mo, ve, va;
        mo.load(am + r * ld, ld, gl); ve.load(av + r * ld, ld, gl);
        float* vr = var.row(r);
        va.load(vr, ld, gl);
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            float* g = &acc.c[q].x; float* mm = &mo.c[q].x; float* vv = &ve.c[q].x; float* xx = &va.c[q].x;
#pragma unroll
            for (int z = 0; z < 4; ++z) {
                mm[z] = 0.9f * mm[z] + 0.1f * g[z];
                vv[z] = 0.999f * vv[z] + 0.001f * g[z] * g[z];
                xx[z] -= lr_t * mm[z] / (sqrtf(vv[z]) + 1e-8f);
            }
        }
        mo.store(am + r * ld, ld, gl); ve.store(av + r * ld, ld, gl); va.store(vr, ld, gl);
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

template <int G, int NC>
__global__ void __launch_bounds__(kGcnThreads, 1) gcn_steps_kernel(const GcnParams p) {
    constexpr int NGRP = kGcnThreads / G;
    __shared__ __align__(16) float part[NGRP * 4 * G * NC];
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
            gcn_product<G, NC>(p, src, epi, part);
            grid.sync();
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
        grid.sync();

        // ---- combine: the first slot of a row sums the row's slots in slot order, norm-backward per layer, stamp ----
        const int ns = 3 * nb;
        for (int64_t sl = ggrp; sl < ns; sl += tot) {
            const int32_t r = __ldcg(p.slot_row + sl);
            bool dup = false;
            for (int64_t x = gl; x < sl; x += G) dup |= __ldcg(p.slot_row + x) == r;
            if (!__any_sync(mask, dup)) {
                GcnVec<G, NC> g; g.load(p.slot_grad + sl * ld, ld, gl);
                for (int64_t x0 = sl + 1; x0 < ns; x0 += G) {
                    const bool hit = x0 + gl < ns && __ldcg(p.slot_row + x0 + gl) == r;
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
        grid.sync();

        // ---- backward: D_k = A D_{k+1} + NB_k, k = L-1 .. 0; k = 0 is the Adam step ----
        const double t_adam = (double)(p.adam_t + s + 1);
        const float lr_t = (float)((double)p.lr * sqrt(1.0 - pow(0.999, t_adam)) / (1.0 - pow(0.9, t_adam)));
        for (int k = p.L - 1; k >= 0; --k) {
            GcnBackwardEpi<G, NC> epi{p.stamp, p.slot_of, p.NB + (int64_t)k * 3 * p.batch * ld, cur, ld,
                                      k > 0 ? p.D[k & 1] : nullptr, gcn_table0(p), p.am, p.av, lr_t};
            if (k == p.L - 1) {
                const GcnStampedSrc<G, NC> src{p.stamp, p.slot_of, p.NB + (int64_t)p.L * 3 * p.batch * ld, cur, ld, p.m};
                gcn_product<G, NC>(p, src, epi, part);
            } else {
                const GcnDenseSrc<G, NC> src{gcn_table(p, p.D[(k + 1) & 1])};
                gcn_product<G, NC>(p, src, epi, part);
            }
            grid.sync();
        }
    }
}

}  // namespace yue
