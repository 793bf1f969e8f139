// K8: CUNE's two-level BPR epoch (replaces the training loop of recommender/advanced/CUNE.py:118-178).
//
// Per training event (u, i) the reference runs THREE repeats (CUNE.py:129).  When the user has implicit positives
// (tracks of its top-K similar users it has not played, CUNE.py:111-113) a repeat draws one of them, k, and one unplayed
// track j, and applies ten row statements (CUNE.py:134-159), each of which RE-EVALUATES its sigmoid on the rows as they
// are at that statement:
//     P[u] += g1 (Q[i] - Q[k]);  Q[i] += g1' P[u];  Q[k] -= g1'' P[u]          g1 = lr (1 - s(P[u].Q[i] - P[u].Q[k]))
//     P[u] += g2 (Q[k] - Q[j]);  Q[k] += g2' P[u];  Q[j] -= g2'' P[u]          g2 = lr/s (1 - s((P[u].Q[k] - P[u].Q[j]) / s))
//     P[u], Q[i], Q[j], Q[k] shrunk by (1 - lr reg)   (in that order; j and k may be the same track: shrunk twice)
//     loss += -log s(x_uik) - log s(x_ukj / s)        on the final rows
// otherwise a plain BPR step on (u, i, j) with re-evaluated sigmoids and no shrink (CUNE.py:164-171).
// After EVERY USER the reference adds regU |P|^2 + regI |Q|^2 over the whole tables to the loss (CUNE.py:174, kept as
// shipped); the serial mode does exactly that, the parallel mode adds (users with events) x the end-of-epoch norms on
// the host side of the C-ABI (documented deviation: the mid-epoch state of a parallel schedule is not defined).
//
// One warp per work item; a lane owns the 16-byte chunks lane, lane+32, ... of every row it touches (128-bit accesses, at
// d = 64 one per row on half the lanes; no lane ever reads a column another lane wrote, so a warp needs no fences between
// its own statements).  P[u] and Q[i]
// stay in registers over the three repeats; both draws are fused (Philox streams of oracle/philox.py: negatives with
// slot n, the implicit positive with slot 64 + n, attempt 0, mapped onto the user's list by (r * len) >> 32).
//   kSerial: one warp walks the users in order, float64 sigmoid and two roundings per axpy like numpy -- the parity anchor.
//   kAtomic: warps take WORK ITEMS from a cursor in stream order; Q changes leave as float adds (Hogwild).  An item is a
//            user, except that a user with more than `chunk` events is cut into items of `chunk` events (on a power-law
//            log the heaviest user alone would otherwise be seconds of one warp's work).  Warps sharing a user publish
//            their change of P[u] as adds and re-read the row every kCuneResync events -- the bound on the staleness of
//            P[u] that K2's quality study found necessary (profiles/quality_study_r1.md) -- everyone else owns P[u].
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#ifndef YUE_CUNE_HOST_EMUL
#include "bpr_sgd.cuh"
#else                    // tests/emul/cune_emul.cpp: this header alone, compiled for the host with a one-lane "warp"
namespace yue { enum : int { kSerial = 0, kAtomic = 1, kStore = 2 }; }
#endif
#include "philox.cuh"

#include <vector>

namespace yue {

constexpr int kCuneResync = 8;
constexpr int64_t kCuneItemWords = 4;      // {user, first event, end event, shared} per work item

// Work items in stream order (host side; the C-ABI and tests/emul use this same text).  chunk <= 0: one item per user.
inline void cune_plan_items(int64_t m, const int64_t* ev_indptr, int64_t chunk, std::vector<int64_t>& items) {
    items.clear();
    for (int64_t u = 0; u < m; ++u) {
        const int64_t b = ev_indptr[u], e = ev_indptr[u + 1];
        if (b == e) continue;                                    // PositiveSet has no entry for a user without events
        if (chunk <= 0 || e - b <= chunk) { items.insert(items.end(), {u, b, e, 0}); continue; }
        for (int64_t x = b; x < e; x += chunk) items.insert(items.end(), {u, x, x + chunk < e ? x + chunk : e, 1});
    }
}

struct CuneParams {
    float* P; float* Q;
    int ld;                          // row stride in floats
    int k;                           // num.factors (columns >= k are padding and stay zero)
    int64_t m, n;
    const int64_t* ev_indptr;        // [m+1] events per user (user-major, BPR.py:42-45 order)
    const int64_t* items;            // [n_work * kCuneItemWords] work items (cune_plan_items)
    int64_t n_work;
    const int32_t* ev_items;         // [T]; v < 0 names hot slot -v-1 (the SGD planner's relabelling)
    const int32_t* hot_items;
    const int64_t* uq_indptr;        // [m+1] sorted-unique play rows (rejection set)
    const int32_t* uq_items;
    const int64_t* ip_indptr;        // [m+1] implicit positives per user
    const int32_t* ip_items;
    uint64_t seed; uint32_t epoch;
    int64_t event_base; const int64_t* ev_delta;
    double lr, inv_s, regU, regI;
    float c_u, c_i;                  // float(lr*regU), float(lr*regI)
    unsigned long long* cursor;
    double* loss;                    // [0] sum of the -log terms (+ the per-user norms in serial mode)
    unsigned long long* users_done;  // users with at least one event (parallel mode: the host scales the norms by it)
    int skip_user_norms;             // serial mode: 1 = leave CUNE.py:174's per-user pass over both tables out of the loss
                                     // (both regs are 0, or a schedule comparison at a size where m passes over the tables
                                     // by one warp would take minutes: YUE_CUNE_EVENT_LOSS=1)
};

constexpr int kCuneMaxC = 2;         // 16-byte chunks per lane: ld <= 256

// W = lanes per warp: 32 on the device; 1 or 8 in the host emulation of tests/emul (same statements).
// A lane owns the 16-byte chunks lane, lane + W, ... of a row (ld % 4 == 0): one 128-bit load per chunk, and the change of
// a shared row leaves as ONE vector add per chunk (red.global.add.v4.f32, like K2) instead of four scalar atomics.
template <int NC, int MODE, int W = 32>
struct CuneRow {
    float4 v[NC];
    static __device__ __forceinline__ bool has(int lane, int c, int ld) { return 4 * (lane + W * c) < ld; }
    __device__ __forceinline__ void load(const float* p, int lane, int ld) {
#pragma unroll
        for (int c = 0; c < NC; ++c) v[c] = has(lane, c, ld) ? ld_row(p + 4 * (lane + W * c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ void store(float* p, int lane, int ld) const {
#pragma unroll
        for (int c = 0; c < NC; ++c) if (has(lane, c, ld)) st_row(p + 4 * (lane + W * c), v[c]);
    }
    // p += (v - old): the change of a shared row leaves as adds
    __device__ __forceinline__ void add_delta(float* p, const CuneRow& old, int lane, int ld) const {
#pragma unroll
        for (int c = 0; c < NC; ++c)
            if (has(lane, c, ld))
                red_row(p + 4 * (lane + W * c), make_float4(v[c].x - old.v[c].x, v[c].y - old.v[c].y, v[c].z - old.v[c].z, v[c].w - old.v[c].w));
    }
    static __device__ __forceinline__ float axpy1(float g, float a, float y) {         // y + g a; serial: two roundings like numpy
        return MODE == kSerial ? __fadd_rn(y, __fmul_rn(g, a)) : fmaf(g, a, y);
    }
    // this += g * (a - b)
    __device__ __forceinline__ void axpy_diff(float g, const CuneRow& a, const CuneRow& b) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            v[c].x = axpy1(g, __fsub_rn(a.v[c].x, b.v[c].x), v[c].x);
            v[c].y = axpy1(g, __fsub_rn(a.v[c].y, b.v[c].y), v[c].y);
            v[c].z = axpy1(g, __fsub_rn(a.v[c].z, b.v[c].z), v[c].z);
            v[c].w = axpy1(g, __fsub_rn(a.v[c].w, b.v[c].w), v[c].w);
        }
    }
    // this += g * a
    __device__ __forceinline__ void axpy(float g, const CuneRow& a) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            v[c].x = axpy1(g, a.v[c].x, v[c].x); v[c].y = axpy1(g, a.v[c].y, v[c].y);
            v[c].z = axpy1(g, a.v[c].z, v[c].z); v[c].w = axpy1(g, a.v[c].w, v[c].w);
        }
    }
    // this -= c * this
    __device__ __forceinline__ void shrink(float cc) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            v[c].x = axpy1(-cc, v[c].x, v[c].x); v[c].y = axpy1(-cc, v[c].y, v[c].y);
            v[c].z = axpy1(-cc, v[c].z, v[c].z); v[c].w = axpy1(-cc, v[c].w, v[c].w);
        }
    }
};

// x = p.a - p.b on every lane
template <int NC, int MODE, int W>
__device__ __forceinline__ float cune_dot_diff(const CuneRow<NC, MODE, W>& p, const CuneRow<NC, MODE, W>& a, const CuneRow<NC, MODE, W>& b) {
    float da = 0.f, db = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        da = fmaf(p.v[c].x, a.v[c].x, da); db = fmaf(p.v[c].x, b.v[c].x, db);
        da = fmaf(p.v[c].y, a.v[c].y, da); db = fmaf(p.v[c].y, b.v[c].y, db);
        da = fmaf(p.v[c].z, a.v[c].z, da); db = fmaf(p.v[c].z, b.v[c].z, db);
        da = fmaf(p.v[c].w, a.v[c].w, da); db = fmaf(p.v[c].w, b.v[c].w, db);
    }
#pragma unroll
    for (int mk = W / 2; mk >= 1; mk >>= 1) {
        da += __shfl_xor_sync(0xffffffffu, da, mk);
        db += __shfl_xor_sync(0xffffffffu, db, mk);
    }
    return da - db;
}

// g = scale * (1 - sigmoid(x)) as float; serial: tool/qmath.py:115-116 in float64 like CPython
template <int MODE>
__device__ __forceinline__ float cune_gain(double scale, float scale_f, float x) {
    if (MODE == kSerial) return (float)(scale * (1.0 - 1.0 / (1.0 + exp(-(double)x))));
    const float ex = __expf(-fabsf(x));
    return scale_f * __fdividef(x >= 0.f ? ex : 1.f, 1.f + ex);    // denominator in [1, 2]: no slow path needed
}
template <int MODE>
__device__ __forceinline__ double cune_nll(float x) {          // -log(sigmoid(x))
    if (MODE == kSerial) return -log(1.0 / (1.0 + exp(-(double)x)));
    return (double)(fmaxf(-x, 0.f) + log1pf(__expf(-fabsf(x))));
}

// sum of squares of a [rows, ld] table by one warp (serial mode's per-user regulariser, CUNE.py:174)
template <int W>
__device__ __forceinline__ double cune_frob2_warp(const float* t, int64_t count, int lane) {
    double s = 0.0;
    for (int64_t x = lane; x < count; x += W) { const float v = __ldcg(t + x); s += (double)v * (double)v; }
#pragma unroll
    for (int mk = W / 2; mk >= 1; mk >>= 1) s += __shfl_xor_sync(0xffffffffu, s, mk);
    return s;
}

template <int NC, int MODE, int W = 32>
__global__ void __launch_bounds__(256, 2) cune_sgd_kernel(const CuneParams p) {
    using Row = CuneRow<NC, MODE, W>;
    const int lane = threadIdx.x & (W - 1);
    const float lr_f = (float)p.lr, lr2_f = (float)(p.inv_s * p.lr), inv_s_f = (float)p.inv_s;
    const double lr2 = p.inv_s * p.lr;
    double loss = 0.0;
    unsigned long long done = 0;

    auto take = [&]() -> int64_t {
        unsigned long long it = 0;
        if (lane == 0) it = atomicAdd(p.cursor, 1ull);
        return (int64_t)__shfl_sync(0xffffffffu, it, 0);
    };
    for (int64_t item = take(); item < p.n_work; item = take()) {
        const int64_t u = p.items[item * kCuneItemWords], eb = p.items[item * kCuneItemWords + 1], ee = p.items[item * kCuneItemWords + 2];
        const bool shared = MODE != kSerial && p.items[item * kCuneItemWords + 3] != 0;
        if (eb == p.ev_indptr[u]) ++done;                        // the user's first item counts the user
        const int64_t r0 = p.uq_indptr[u];
        const int32_t* row = p.uq_items + r0;
        const int row_len = (int)(p.uq_indptr[u + 1] - r0);
        const int64_t ip0 = p.ip_indptr[u];
        const uint32_t ip_len = (uint32_t)(p.ip_indptr[u + 1] - ip0);
        const int64_t gbase = p.ev_delta ? p.ev_delta[u] : p.event_base;
        float* const pu_ptr = p.P + (size_t)u * p.ld;
        Row pu, pu0;
        pu.load(pu_ptr, lane, p.ld);
        pu0 = pu;

        for (int64_t e0 = eb; e0 < ee; e0 += W) {
            const int len = (int)((ee - e0) < W ? (ee - e0) : W);
            // lane t draws for event e0 + t: the positive, three negatives, three implicit positives
            int32_t my_i = 0, my_j[3] = {0, 0, 0}, my_k[3] = {0, 0, 0};
            if (lane < len) {
                const int64_t e = e0 + lane;
                const uint64_t ge = (uint64_t)(gbase + e);
                my_i = p.ev_items[e];
                if (my_i < 0) my_i = p.hot_items[-my_i - 1];
#pragma unroll
                for (int nn = 0; nn < 3; ++nn) {
                    my_j[nn] = sample_negative(p.seed, p.epoch, ge, (uint32_t)nn, (uint32_t)p.n, row, row_len);
                    if (ip_len > 0) {
                        uint32_t w[4];
                        philox4x32_10((uint32_t)ge, (uint32_t)(ge >> 32), p.epoch, (uint32_t)(64 + nn) << 20,
                                      (uint32_t)p.seed, (uint32_t)(p.seed >> 32), w);
                        my_k[nn] = p.ip_items[ip0 + (int64_t)__umulhi(w[0], ip_len)];
                    }
                }
            }
            __syncwarp();
            for (int t = 0; t < len; ++t) {
                if (shared && (e0 - eb + t) > 0 && (e0 - eb + t) % kCuneResync == 0) {
                    pu.add_delta(pu_ptr, pu0, lane, p.ld);       // publish, then re-read: same thread, same address -> ordered
                    pu.load(pu_ptr, lane, p.ld);
                    pu0 = pu;
                }
                const int32_t i = __shfl_sync(0xffffffffu, my_i, t);
                float* const qi_ptr = p.Q + (size_t)i * p.ld;
                Row qi, qi0;
                qi.load(qi_ptr, lane, p.ld);
                qi0 = qi;
#pragma unroll
                for (int nn = 0; nn < 3; ++nn) {
                    // selects, not my_j[nn]: the arrays stay in registers whether or not the loop is unrolled
                    const int32_t j = __shfl_sync(0xffffffffu, nn == 0 ? my_j[0] : (nn == 1 ? my_j[1] : my_j[2]), t);
                    const int32_t k = __shfl_sync(0xffffffffu, nn == 0 ? my_k[0] : (nn == 1 ? my_k[1] : my_k[2]), t);
                    if (j < 0) continue;                         // the user has played the whole catalog: no negative exists
                    float* const qj_ptr = p.Q + (size_t)j * p.ld;
                    Row qj, qj0;
                    qj.load(qj_ptr, lane, p.ld);
                    qj0 = qj;
                    if (ip_len > 0) {
                        float* const qk_ptr = p.Q + (size_t)k * p.ld;
                        const bool alias = k == j;               // both unplayed by u: may be the same track
                        Row qk, qk0;
                        if (alias) qk = qj; else qk.load(qk_ptr, lane, p.ld);
                        qk0 = qk;
                        float g;
                        g = cune_gain<MODE>(p.lr, lr_f, cune_dot_diff(pu, qi, qk));  pu.axpy_diff(g, qi, qk);   // 134-135
                        g = cune_gain<MODE>(p.lr, lr_f, cune_dot_diff(pu, qi, qk));  qi.axpy(g, pu);            // 136-137
                        g = cune_gain<MODE>(p.lr, lr_f, cune_dot_diff(pu, qi, qk));  qk.axpy(-g, pu);           // 138-139
                        if (alias) qj = qk;
                        g = cune_gain<MODE>(lr2, lr2_f, inv_s_f * cune_dot_diff(pu, qk, qj));  pu.axpy_diff(g, qk, qj);   // 148-149
                        g = cune_gain<MODE>(lr2, lr2_f, inv_s_f * cune_dot_diff(pu, qk, qj));  qk.axpy(g, pu);            // 150-151
                        if (alias) qj = qk;
                        g = cune_gain<MODE>(lr2, lr2_f, inv_s_f * cune_dot_diff(pu, qk, qj));  qj.axpy(-g, pu);           // 152-154
                        if (alias) qk = qj;
                        pu.shrink(p.c_u);                                                                       // 156
                        qi.shrink(p.c_i);                                                                       // 157
                        qj.shrink(p.c_i);                                                                       // 158
                        if (alias) qk = qj;
                        qk.shrink(p.c_i);                                                                       // 159
                        if (alias) qj = qk;
                        loss += cune_nll<MODE>(cune_dot_diff(pu, qi, qk)) + cune_nll<MODE>(inv_s_f * cune_dot_diff(pu, qk, qj));   // 161-162
                        if (!alias) { if (MODE == kAtomic) qk.add_delta(qk_ptr, qk0, lane, p.ld); else qk.store(qk_ptr, lane, p.ld); }
                    } else {
                        float g;
                        g = cune_gain<MODE>(p.lr, lr_f, cune_dot_diff(pu, qi, qj));  pu.axpy_diff(g, qi, qj);   // 164-165
                        g = cune_gain<MODE>(p.lr, lr_f, cune_dot_diff(pu, qi, qj));  qi.axpy(g, pu);            // 166-167
                        g = cune_gain<MODE>(p.lr, lr_f, cune_dot_diff(pu, qi, qj));  qj.axpy(-g, pu);           // 168-169
                        loss += cune_nll<MODE>(cune_dot_diff(pu, qi, qj));                                      // 171
                    }
                    if (MODE == kAtomic) qj.add_delta(qj_ptr, qj0, lane, p.ld); else qj.store(qj_ptr, lane, p.ld);
                }
                if (MODE == kAtomic) qi.add_delta(qi_ptr, qi0, lane, p.ld); else qi.store(qi_ptr, lane, p.ld);
            }
        }
        if (shared) pu.add_delta(pu_ptr, pu0, lane, p.ld);
        else pu.store(pu_ptr, lane, p.ld);                       // P[u] belongs to this warp alone
        if (MODE == kSerial && !p.skip_user_norms) {             // CUNE.py:174, once per user, over the whole tables
            __syncwarp();                                        // the sums read columns other lanes stored (ld % 32 != 0)
            loss += p.regU * cune_frob2_warp<W>(p.P, p.m * p.ld, lane) + p.regI * cune_frob2_warp<W>(p.Q, p.n * p.ld, lane);
        }
    }
    if (lane == 0) {
        if (loss != 0.0) atomicAdd(p.loss, loss);
        if (done) atomicAdd(p.users_done, done);
    }
}

}  // namespace yue
