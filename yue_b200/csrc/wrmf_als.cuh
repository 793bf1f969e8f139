// K7: WRMF half-sweep (replaces the two loops of recommender/cf/WRMF.py:34-80).
//
// The reference solves, for every user u (and then, roles swapped, for every track),
//     A = YtY + Y^T diag(10 r_ui) Y + reg I,   b = sum_i (1 + 10 r_ui) Y[i],   X[u] = inv(A) b
// with an n-long dense H vector, an n x n scipy COO matrix and a dense inverse per row.  Rows of a sweep are
// independent (a solve reads only the OTHER table), so a sweep here is:
//   wrmf_gram_kernel     G = F^T F of the other table, float64 accumulation of the float32 rows, per-CTA partials
//                        summed in a fixed order (wrmf_gram_reduce_kernel) -- same bits on every call;
//   wrmf_chunk_kernel    rows with more than kWrmfChunk entries (the heaviest users, the most played tracks: one
//                        track of config C2 has ~6e5 listeners) are cut into chunks whose weighted Gram + rhs partials
//                        go to scratch, so that no single CTA is a straggler;
//   wrmf_binv_kernel +   rows with 1..16 entries (half the users of config C2): one warp per row on the d x d Woodbury
//   wrmf_light_kernel    system with B^-1 = (G + reg I)^-1 computed once per sweep (see "light rows" below);
//   wrmf_solve_kernel    every other row, one CTA per row (rows strided over the grid, the next row's data prefetched):
//                        A = G + reg I + sum of the row's rank-1 terms (or of its chunk partials), LDL^T in registers,
//                        forward substitution carried as an extra row of the factorisation, backward substitution by
//                        row blocks, X[row] stored as float32.
//
// Tile layout.  k is padded to KP = 16 TD (TD = 1, 2, 4, 8 for k <= 16, 32, 64, 128).  A is symmetric: only the 136
// lower-triangular TD x TD blocks of the 16 x 16 block grid exist, one per thread (160 threads, 24 of them only help
// with loads), each block in REGISTERS from the first rank-1 term to the last substitution step.  The only matrix data
// that ever passes through shared memory is the current column of the factorisation (double-buffered: one
// __syncthreads per column).  All arithmetic is float64 like the reference's (the sparse int64 weights promote its
// products to float64, WRMF.py:52-54); float64 FMA runs at half the float32 rate on sm_100, and the kernel's floor is
// the rank-1 accumulation: nnz * KP^2 / 2 FMAs per sweep.
//
// Loss (WRMF.py:49-50): sum (1 - X[u].Y[i])^2 over the played pairs with the X[u] from before its update, computed by
// the warps on the staged rows while they are in shared memory.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace yue {

// Shape of the tiling: A (k padded to KP = TD * GB) is a GB x GB grid of TD x TD blocks, of which the NB = GB (GB + 1) / 2
// lower-triangular ones exist, one per thread.  <TD, 16>: 136 blocks on 160 threads (k <= 16 TD), the default; <8, 8>: 36
// blocks on 64 threads for k <= 64 -- fewer, fatter threads: 64 FMAs per thread and column step instead of 16, a barrier
// between 2 warps instead of 5, half the shared-memory traffic per step -- measured slower (YUE_WRMF_FAT=1 selects it).
template <int TD, int GB>
struct WrmfShape {
    static constexpr int KP = TD * GB;
    static constexpr int NB = GB * (GB + 1) / 2;
    static constexpr int NT = (NB + 31) / 32 * 32;
};
constexpr int kWrmfBatch = 16;           // rows of the other table staged per step
constexpr int kWrmfChunk = 4096;         // entries one CTA accumulates for a row at most

struct WrmfSide {
    float* out;                 // table being solved [rows, ld]
    const float* other;         // the fixed table [*, ld]
    const int64_t* indptr;      // [rows + 1]
    const int32_t* idx;         // rows of `other` per entry
    const int32_t* cnt;         // plays per entry (r_ui)
    int64_t row_begin, rows;    // this call solves rows [row_begin, rows)
    int chunk_off;              // first chunk of the heavy rows in that range
    int light_max;              // rows with 1..light_max entries are solved by wrmf_light_kernel (0: none)
    int ld, k;
    double reg, alpha;
    const double* G;            // Gram matrix of `other`, thread layout [TD*TD][kWrmfThreads]
    // heavy rows
    const int32_t* heavy_rows;  // sorted row ids with more than kWrmfChunk entries
    const int32_t* heavy_first; // [n_heavy + 1] first chunk of each
    int n_heavy;
    const int64_t* chunk_begin; // [n_chunks] entry range of every chunk
    const int64_t* chunk_end;
    const int32_t* chunk_row;
    double* partA;              // [n_chunks][TD*TD][kWrmfThreads]
    double* partb;              // [n_chunks][KP]
    double* loss;               // += sum of squared errors, or nullptr
};

template <int TD>
struct WrmfTile {
    double acc[TD][TD];         // rows ty*TD + i, columns tx*TD + j
    double bb[TD];              // rhs of columns tx*TD + j (diagonal threads only)
};

// 1/d for a positive, normal d: float32 reciprocal as the seed, two Newton steps in float64 (relative error ~2^-46
// after the first, below 2^-52 after the second).  ~4 dependent FMAs instead of the ~25-instruction IEEE division;
// the pivots of G + reg I + (rank-1 terms) are positive and far from the float32 range limits.
__device__ __forceinline__ double wrmf_rcp(double d) {
    double x = (double)__frcp_rn((float)d);
    x = fma(x, fma(-d, x, 1.0), x);
    x = fma(x, fma(-d, x, 1.0), x);
    return x;
}

// Thread -> block of A, column by column: threads 0..15 own block column 0 (block rows 0..15), threads 16..30 column 1
// (rows 1..15), ...  A warp then shares (almost) one block column: its reads of the column's entries are broadcasts,
// its reads of the row entries are contiguous, and once the factorisation has passed a warp's columns the whole warp
// skips the update -- the shared-memory wavefronts and the float64 issue slots of a step shrink with the trailing matrix.
template <int GB>
__device__ __forceinline__ void wrmf_block_of_thread(int t, int& ty, int& tx) {
    int c = 0, start = 0;
    while (start + (GB - c) <= t) { start += GB - c; ++c; }
    tx = c; ty = c + (t - start);
}

// TD consecutive doubles from 16-byte aligned shared memory (LDS.128 for TD >= 2)
template <int TD>
__device__ __forceinline__ void wrmf_lds(const double* p, double (&v)[TD]) {
    if constexpr (TD == 1) {
        v[0] = p[0];
    } else {
#pragma unroll
        for (int i = 0; i < TD; i += 2) {
            const double2 t = *reinterpret_cast<const double2*>(p + i);
            v[i] = t.x; v[i + 1] = t.y;
        }
    }
}

// One staged batch in flight: kWrmfBatch rows of the other table (this thread's float4 slots) and one weight.
template <int TD, int GB>
struct WrmfPre {
    static constexpr int Q4 = TD * GB / 4;                       // float4 per staged row
    static constexpr int NLOAD = (kWrmfBatch * Q4 + WrmfShape<TD, GB>::NT - 1) / WrmfShape<TD, GB>::NT;
    float4 v[NLOAD];
    double w;
};

// issue the loads of entries [eb, min(eb + kWrmfBatch, e1)): rows idx[e] of `other` (rows e themselves when idx ==
// nullptr) and their weights alpha * cnt[e] (1 when cnt == nullptr)
template <int TD, int GB>
__device__ __forceinline__ void wrmf_fetch(WrmfPre<TD, GB>& pre, const float* __restrict__ other, int ld, const int32_t* __restrict__ idx,
                                           const int32_t* __restrict__ cnt, double alpha, int64_t eb, int64_t e1) {
    constexpr int Q4 = WrmfPre<TD, GB>::Q4, kWrmfThreads = WrmfShape<TD, GB>::NT;
    const int tid = threadIdx.x;
#pragma unroll
    for (int q = 0; q < WrmfPre<TD, GB>::NLOAD; ++q) {
        const int s = tid + q * kWrmfThreads;
        const int r = s / Q4, c4 = s % Q4;
        pre.v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < kWrmfBatch && eb + r < e1 && c4 * 4 < ld) {
            const int64_t row = idx ? (int64_t)idx[eb + r] : eb + r;
            pre.v[q] = __ldg(reinterpret_cast<const float4*>(other + row * ld) + c4);
        }
    }
    pre.w = 0.0;
    if (tid < kWrmfBatch && eb + tid < e1) pre.w = cnt ? alpha * (double)cnt[eb + tid] : 1.0;
}

// acc += sum_e w_e y_e y_e^T, bb += sum_e (1 + w_e) y_e over entries [e0, e1) (bb only when cnt != nullptr).  `pre` holds
// the first batch already (wrmf_fetch(pre, ..., e0, e1) issued by the caller, possibly long ago).  ys: smem
// [kWrmfBatch][KP] doubles, ws: smem [kWrmfBatch], xs: smem [KP] (the row's current solution, LOSS only).
template <int TD, int GB, bool LOSS>
__device__ __forceinline__ void wrmf_accumulate(WrmfTile<TD>& tl, WrmfPre<TD, GB>& pre, const float* __restrict__ other, int ld,
                                                const int32_t* __restrict__ idx, const int32_t* __restrict__ cnt, int64_t e0,
                                                int64_t e1, double alpha, double* ys, double* ws, const double* xs, double& loss,
                                                int ty, int tx, bool active) {
    constexpr int KP = TD * GB, kWrmfThreads = WrmfShape<TD, GB>::NT;
    constexpr int Q4 = WrmfPre<TD, GB>::Q4;
    const int tid = threadIdx.x;
    for (int64_t eb = e0; eb < e1; eb += kWrmfBatch) {
#pragma unroll
        for (int q = 0; q < WrmfPre<TD, GB>::NLOAD; ++q) {
            const int s = tid + q * kWrmfThreads;
            const int r = s / Q4, c4 = s % Q4;
            if (r < kWrmfBatch) {
                double* d = ys + r * KP + c4 * 4;
                // columns >= k of a padded table row are zero by construction (yue_set_factors)
                d[0] = pre.v[q].x; d[1] = pre.v[q].y; d[2] = pre.v[q].z; d[3] = pre.v[q].w;
            }
        }
        if (tid < kWrmfBatch) ws[tid] = pre.w;
        __syncthreads();
        if (eb + kWrmfBatch < e1) wrmf_fetch<TD, GB>(pre, other, ld, idx, cnt, alpha, eb + kWrmfBatch, e1);   // in flight under the FMAs
        const int nb = (int)((e1 - eb) < (int64_t)kWrmfBatch ? (e1 - eb) : (int64_t)kWrmfBatch);
        if (active) {
            for (int e = 0; e < nb; ++e) {
                const double* y = ys + e * KP;
                const double w = ws[e];
                double a[TD], b[TD];
                wrmf_lds<TD>(y + ty * TD, a);
                wrmf_lds<TD>(y + tx * TD, b);
#pragma unroll
                for (int i = 0; i < TD; ++i) a[i] *= w;
#pragma unroll
                for (int i = 0; i < TD; ++i)
#pragma unroll
                    for (int j = 0; j < TD; ++j) tl.acc[i][j] = fma(a[i], b[j], tl.acc[i][j]);
                if (cnt != nullptr && ty == tx) {
#pragma unroll
                    for (int j = 0; j < TD; ++j) tl.bb[j] = fma(1.0 + w, b[j], tl.bb[j]);
                }
            }
        }
        if (LOSS) {
            const int lane = tid & 31, warp = tid >> 5;
            for (int e = warp; e < nb; e += kWrmfThreads / 32) {
                double p = 0.0;
                for (int c = lane; c < KP; c += 32) p = fma(xs[c], ys[e * KP + c], p);
#pragma unroll
                for (int s = 16; s >= 1; s >>= 1) p += __shfl_xor_sync(0xffffffffu, p, s);
                if (lane == 0) { const double err = 1.0 - p; loss = fma(err, err, loss); }
            }
        }
        __syncthreads();
    }
}

// G partial of rows [row0, row1) of F (unit weights), thread layout
template <int TD, int GB>
__global__ void __launch_bounds__(WrmfShape<TD, GB>::NT) wrmf_gram_kernel(const float* __restrict__ F, int64_t n, int ld, int k,
                                                                 double* __restrict__ partial) {
    constexpr int KP = TD * GB, kWrmfThreads = WrmfShape<TD, GB>::NT, kWrmfBlocks = WrmfShape<TD, GB>::NB;
    extern __shared__ __align__(16) double wrmf_smem[];
    double* ys = wrmf_smem;
    double* ws = ys + kWrmfBatch * KP;
    const int tid = threadIdx.x;
    int ty = 0, tx = 0;
    bool active = tid < kWrmfBlocks;
    if (active) wrmf_block_of_thread<GB>(tid, ty, tx);
    active = active && ty < (k + TD - 1) / TD;           // blocks of padded unknowns only: nothing to do
    WrmfTile<TD> tl;
#pragma unroll
    for (int i = 0; i < TD; ++i) { tl.bb[i] = 0.0;
#pragma unroll
        for (int j = 0; j < TD; ++j) tl.acc[i][j] = 0.0; }
    const int64_t r0 = n * (int64_t)blockIdx.x / gridDim.x, r1 = n * (int64_t)(blockIdx.x + 1) / gridDim.x;
    double dummy = 0.0;
    WrmfPre<TD, GB> pre;
    wrmf_fetch<TD, GB>(pre, F, ld, nullptr, nullptr, 1.0, r0, r1);
    wrmf_accumulate<TD, GB, false>(tl, pre, F, ld, nullptr, nullptr, r0, r1, 1.0, ys, ws, nullptr, dummy, ty, tx, active);
    double* out = partial + (size_t)blockIdx.x * TD * TD * kWrmfThreads;
#pragma unroll
    for (int i = 0; i < TD; ++i)
#pragma unroll
        for (int j = 0; j < TD; ++j) out[(i * TD + j) * kWrmfThreads + tid] = active ? tl.acc[i][j] : 0.0;
}

// G[e] = sum over the partials in index order
__global__ void wrmf_gram_reduce_kernel(const double* __restrict__ partial, int n_part, int elems, double* __restrict__ G) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= elems) return;
    double s = 0.0;
    for (int p = 0; p < n_part; ++p) s += partial[(size_t)p * elems + e];
    G[e] = s;
}

template <int TD, int GB, bool LOSS>
__global__ void __launch_bounds__(WrmfShape<TD, GB>::NT) wrmf_chunk_kernel(WrmfSide sd) {
    constexpr int KP = TD * GB, kWrmfThreads = WrmfShape<TD, GB>::NT, kWrmfBlocks = WrmfShape<TD, GB>::NB;
    extern __shared__ __align__(16) double wrmf_smem[];
    double* ys = wrmf_smem;
    double* ws = ys + kWrmfBatch * KP;
    double* xs = ws + kWrmfBatch;
    const int tid = threadIdx.x;
    int ty = 0, tx = 0;
    bool active = tid < kWrmfBlocks;
    if (active) wrmf_block_of_thread<GB>(tid, ty, tx);
    active = active && ty < (sd.k + TD - 1) / TD;
    const int ch = blockIdx.x + sd.chunk_off;
    const int64_t row = sd.chunk_row[ch];
    if (LOSS) for (int c = tid; c < KP; c += kWrmfThreads) xs[c] = c < sd.k ? (double)sd.out[row * sd.ld + c] : 0.0;
    WrmfTile<TD> tl;
#pragma unroll
    for (int i = 0; i < TD; ++i) { tl.bb[i] = 0.0;
#pragma unroll
        for (int j = 0; j < TD; ++j) tl.acc[i][j] = 0.0; }
    double loss = 0.0;
    const int64_t c0 = sd.chunk_begin[ch], c1 = sd.chunk_end[ch];
    WrmfPre<TD, GB> pre;
    wrmf_fetch<TD, GB>(pre, sd.other, sd.ld, sd.idx, sd.cnt, sd.alpha, c0, c1);
    wrmf_accumulate<TD, GB, LOSS>(tl, pre, sd.other, sd.ld, sd.idx, sd.cnt, c0, c1, sd.alpha, ys, ws, xs, loss, ty, tx, active);
    double* pa = sd.partA + (size_t)ch * TD * TD * kWrmfThreads;
#pragma unroll
    for (int i = 0; i < TD; ++i)
#pragma unroll
        for (int j = 0; j < TD; ++j) pa[(i * TD + j) * kWrmfThreads + tid] = active ? tl.acc[i][j] : 0.0;
    if (active && ty == tx) {
#pragma unroll
        for (int j = 0; j < TD; ++j) sd.partb[(size_t)ch * KP + tx * TD + j] = tl.bb[j];
    }
    if (LOSS) {
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, s);
        if ((tid & 31) == 0 && loss != 0.0) atomicAdd(sd.loss, loss);
    }
}

template <int TD, int GB, bool LOSS>
__global__ void __launch_bounds__(WrmfShape<TD, GB>::NT, (TD * GB <= 64 ? 4 : 1)) wrmf_solve_kernel(WrmfSide sd) {
    constexpr int KP = TD * GB, kWrmfThreads = WrmfShape<TD, GB>::NT, kWrmfBlocks = WrmfShape<TD, GB>::NB;
    extern __shared__ __align__(16) double wrmf_smem[];
    double* ys = wrmf_smem;                              // [kWrmfBatch][KP]
    double* ws = ys + kWrmfBatch * KP;                   // [kWrmfBatch]
    double* xs = ws + kWrmfBatch;                        // [KP]  solution before the update (loss)
    double* col = xs + KP;                               // [2][KP + 2]  current column of the factorisation, the rhs row, 1/pivot
    double* dv = col + 2 * (KP + 2);                     // [KP]  reciprocal pivots
    double* yv = dv + KP;                                // [KP]  D^-1 L^-1 b, then consumed by the back substitution
    double* xv = yv + KP;                                // [KP]  the solution
    double* gs = xv + KP;                                // [TD*TD][kWrmfThreads]  G + reg I
    const int tid = threadIdx.x;
    int ty = 0, tx = 0;
    const int k = sd.k, ld = sd.ld;
    const int nblk = (k + TD - 1) / TD;                  // block rows / columns that hold real unknowns
    bool active = tid < kWrmfBlocks;
    if (active) wrmf_block_of_thread<GB>(tid, ty, tx);
    active = active && ty < nblk;                        // blocks of padded unknowns only: nothing to do
    const bool diag = active && ty == tx;
#pragma unroll
    for (int i = 0; i < TD; ++i)
#pragma unroll
        for (int j = 0; j < TD; ++j) {
            double g = sd.G[(i * TD + j) * kWrmfThreads + tid];
            if (diag && i == j) g = (ty * TD + i < k) ? g + sd.reg : 1.0;      // padded unknowns: identity
            gs[(i * TD + j) * kWrmfThreads + tid] = g;
        }
    double loss = 0.0;
    // Rows blockIdx.x, blockIdx.x + gridDim.x, ...  While a row is being factorised (~64 dependent steps) the first batch
    // of the NEXT row -- entry range, row ids, rows of the other table: three dependent global loads -- is already in
    // flight into registers, so a row starts with its data on chip.
    int64_t row = sd.row_begin + blockIdx.x;
    int64_t e0 = 0, e1 = 0;
    WrmfPre<TD, GB> pre;
    if (row < sd.rows) {
        e0 = sd.indptr[row]; e1 = sd.indptr[row + 1];
        if (e1 - e0 > sd.light_max && e1 - e0 <= kWrmfChunk) wrmf_fetch<TD, GB>(pre, sd.other, ld, sd.idx, sd.cnt, sd.alpha, e0, e1);
    }
    for (; row < sd.rows; ) {
        __syncthreads();                                  // smem of the previous row is free
        const int64_t nrow = row + gridDim.x;
        int64_t ne0 = 0, ne1 = 0;
        if (nrow < sd.rows) { ne0 = sd.indptr[nrow]; ne1 = sd.indptr[nrow + 1]; }
        if (e1 - e0 <= sd.light_max) {                    // a light row (wrmf_light_kernel's), or an empty one:
            if (e1 == e0)                                 // nobody played it / played nothing: b = 0 -> the row is 0
                for (int c = tid; c < ld; c += kWrmfThreads) sd.out[row * ld + c] = 0.f;
            row = nrow; e0 = ne0; e1 = ne1;
            if (row < sd.rows && e1 - e0 > sd.light_max && e1 - e0 <= kWrmfChunk)
                wrmf_fetch<TD, GB>(pre, sd.other, ld, sd.idx, sd.cnt, sd.alpha, e0, e1);
            continue;
        }
        if (LOSS) for (int c = tid; c < KP; c += kWrmfThreads) xs[c] = c < k ? (double)sd.out[row * ld + c] : 0.0;
        WrmfTile<TD> tl;
#pragma unroll
        for (int i = 0; i < TD; ++i) { tl.bb[i] = 0.0;
#pragma unroll
            for (int j = 0; j < TD; ++j) tl.acc[i][j] = gs[(i * TD + j) * kWrmfThreads + tid]; }
        if (e1 - e0 <= kWrmfChunk) {
            wrmf_accumulate<TD, GB, LOSS>(tl, pre, sd.other, ld, sd.idx, sd.cnt, e0, e1, sd.alpha, ys, ws, xs, loss, ty, tx, active);
        } else {                                          // chunk partials, in chunk order
            int lo = 0, hi = sd.n_heavy;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (sd.heavy_rows[mid] < row) lo = mid + 1; else hi = mid; }
            for (int ch = sd.heavy_first[lo]; ch < sd.heavy_first[lo + 1]; ++ch) {
                const double* pa = sd.partA + (size_t)ch * TD * TD * kWrmfThreads;
#pragma unroll
                for (int i = 0; i < TD; ++i)
#pragma unroll
                    for (int j = 0; j < TD; ++j) tl.acc[i][j] += pa[(i * TD + j) * kWrmfThreads + tid];
                if (diag) {
#pragma unroll
                    for (int j = 0; j < TD; ++j) tl.bb[j] += sd.partb[(size_t)ch * KP + tx * TD + j];
                }
            }
        }
        if (nrow < sd.rows && ne1 - ne0 > sd.light_max && ne1 - ne0 <= kWrmfChunk)
            wrmf_fetch<TD, GB>(pre, sd.other, ld, sd.idx, sd.cnt, sd.alpha, ne0, ne1);
        // ---- LDL^T, one column per step; the rhs rides along as row KP of the matrix.  The pivot's reciprocal is
        //      computed once, by the thread that owns the pivot, and published with the column (slot KP + 1).
        //      (A variant that works by block columns of TD unknowns -- two barriers per TD columns, TD^3 FMAs per
        //      2 TD^2 shared-memory words -- was measured SLOWER, 76 vs 65 ms per user sweep at config C2: the serial
        //      factorisation of the diagonal block by one thread stretches the chain the barriers wait on.) ----
        for (int jb = 0; jb < nblk; ++jb) {
#pragma unroll
            for (int jj = 0; jj < TD; ++jj) {
                const int j = jb * TD + jj;
                double* cb = col + (j & 1) * (KP + 2);
                if (active && tx == jb) {
#pragma unroll
                    for (int i = 0; i < TD; ++i) cb[ty * TD + i] = tl.acc[i][jj];
                    if (ty == jb) {
                        const double inv = wrmf_rcp(tl.acc[jj][jj]);
                        cb[KP] = tl.bb[jj]; cb[KP + 1] = inv; dv[j] = inv;
                    }
                }
                __syncthreads();
                if (active && tx >= jb) {
                    const double inv = cb[KP + 1];
                    double rv[TD];
                    wrmf_lds<TD>(cb + ty * TD, rv);
                    const double rb = cb[KP];
#pragma unroll
                    for (int jc = 0; jc < TD; ++jc) {
                        if (tx > jb || jc > jj) {
                            const double cv = cb[tx * TD + jc] * inv;
#pragma unroll
                            for (int i = 0; i < TD; ++i) tl.acc[i][jc] = fma(-rv[i], cv, tl.acc[i][jc]);
                            if (ty == tx) tl.bb[jc] = fma(-rb, cv, tl.bb[jc]);
                        }
                    }
                }
            }
        }
        // ---- y = D^-1 z (z = the forward-substituted rhs now in bb), then x = L^-T y by row blocks, last block first;
        //      dv holds the reciprocal pivots, so no division is left ----
        if (diag) {
#pragma unroll
            for (int j = 0; j < TD; ++j) yv[tx * TD + j] = tl.bb[j] * dv[tx * TD + j];
        }
        __syncthreads();
        for (int rbk = nblk - 1; rbk >= 0; --rbk) {
            if (diag && ty == rbk) {
                double xl[TD];
#pragma unroll
                for (int i = TD - 1; i >= 0; --i) {
                    double s = 0.0;
#pragma unroll
                    for (int i2 = TD - 1; i2 > i; --i2) s = fma(tl.acc[i2][i], xl[i2], s);
                    xl[i] = fma(-s, dv[rbk * TD + i], yv[rbk * TD + i]);
                    xv[rbk * TD + i] = xl[i];
                }
            }
            __syncthreads();
            if (active && ty == rbk && tx < rbk) {
#pragma unroll
                for (int jc = 0; jc < TD; ++jc) {
                    double s = 0.0;
#pragma unroll
                    for (int i = 0; i < TD; ++i) s = fma(tl.acc[i][jc], xv[rbk * TD + i], s);
                    yv[tx * TD + jc] = fma(-s, dv[tx * TD + jc], yv[tx * TD + jc]);
                }
            }
            __syncthreads();
        }
        for (int c = tid; c < k; c += kWrmfThreads) sd.out[row * ld + c] = (float)xv[c];
        row = nrow; e0 = ne0; e1 = ne1;
    }
    if (LOSS) {
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, s);
        if ((tid & 31) == 0 && loss != 0.0) atomicAdd(sd.loss, loss);
    }
}

template <int TD, int GB>
constexpr size_t wrmf_solve_smem() {
    constexpr int KP = TD * GB, kWrmfThreads = WrmfShape<TD, GB>::NT;
    return sizeof(double) * (size_t)(kWrmfBatch * KP + kWrmfBatch + KP + 2 * (KP + 2) + 3 * KP + TD * TD * kWrmfThreads);
}
template <int TD, int GB>
constexpr size_t wrmf_accum_smem() {
    constexpr int KP = TD * GB;
    return sizeof(double) * (size_t)(kWrmfBatch * KP + kWrmfBatch + KP);
}

// ---- light rows: 1..16 entries, k <= 64 ---------------------------------------------------------------------------
// Half the rows of a play log are short (config C2: 48 % of the users have at most 16 tracks, 78 % at most 32).  For them
// the k x k factorisation above is the wrong shape of work: ~64 dependent steps with a block-wide barrier each, for a
// matrix that is a rank-d update (d = entries of the row) of one matrix every row shares, B = G + reg I.  With Y_u the d
// rows of the other table, C = diag(alpha r) and p = 1 + alpha r (Woodbury):
//     x = (B + Y_u^T C Y_u)^-1 Y_u^T p = B^-1 Y_u^T t,    (C^-1 + Y_u B^-1 Y_u^T) t = C^-1 p
// i.e. a d x d system instead of a k x k one, and B^-1 is computed once per sweep (wrmf_binv_kernel).  One WARP per
// row, no block-wide barrier, everything a row needs in ~13 KB of shared memory (14 rows in flight per SM):
//   Z = Y_u B^-1: lane (h, c) owns columns c, 16 + c, 32 + c, 48 + c for the entries e = 2i + h -- all 32 lanes busy
//       whatever d is, every value of B^-1 and of Y_u read once per row (the y values as half-warp broadcasts);
//   M += Z_chunk Y_u,chunk^T: lane (e, h) owns row e, columns 8h .. 8h+7 of M, Z_chunk passed through shared memory;
//   S = M + C^-1 eliminated in shared memory (run-time loops: small code), back substitution, q = Y_u^T t and
//       x = B^-1 q with lane j owning columns j and j + 32.
// All float64, no atomics except the loss: a row gives the same bits in any launch shape.
constexpr int kWrmfLightMax = 16;        // entries of a light row at most
constexpr int kWrmfLightWarps = 14;      // warps (rows in flight) per CTA, one CTA per SM

// B^-1 by in-place Gauss-Jordan (B is symmetric positive definite: no pivoting), one CTA.  G in thread layout.
template <int TD, int GB>
__global__ void __launch_bounds__(256) wrmf_binv_kernel(const double* __restrict__ G, int k, double reg, double* __restrict__ Binv) {
    constexpr int KP = TD * GB, kWrmfThreads = WrmfShape<TD, GB>::NT, kWrmfBlocks = WrmfShape<TD, GB>::NB;
    extern __shared__ __align__(16) double wrmf_smem[];
    double* A = wrmf_smem;                 // [KP][KP + 1]
    double* colp = A + KP * (KP + 1);      // [KP]
    double* rowp = colp + KP;              // [KP]
    const int tid = threadIdx.x;
    if (tid < kWrmfBlocks) {
        int ty, tx;
        wrmf_block_of_thread<GB>(tid, ty, tx);
#pragma unroll
        for (int i = 0; i < TD; ++i)
#pragma unroll
            for (int j = 0; j < TD; ++j) {
                const int r = ty * TD + i, c = tx * TD + j;
                double g = (r < k && c < k) ? G[(i * TD + j) * kWrmfThreads + tid] : 0.0;
                if (r == c) g = r < k ? g + reg : 1.0;
                A[r * (KP + 1) + c] = g;
                A[c * (KP + 1) + r] = g;
            }
    }
    for (int p = 0; p < KP; ++p) {
        __syncthreads();
        if (tid < KP) { colp[tid] = A[tid * (KP + 1) + p]; rowp[tid] = A[p * (KP + 1) + tid]; }
        __syncthreads();
        const double inv = 1.0 / colp[p];
        for (int e = tid; e < KP * KP; e += 256) {
            const int r = e / KP, c = e % KP;
            double* a = A + r * (KP + 1) + c;
            if (r == p) *a = (c == p) ? inv : rowp[c] * inv;
            else if (c == p) *a = -colp[r] * inv;
            else *a = fma(-colp[r] * inv, rowp[c], *a);
        }
    }
    __syncthreads();
    for (int e = tid; e < KP * KP; e += 256) Binv[e] = A[(e / KP) * (KP + 1) + e % KP];
}

template <int KP>
__host__ __device__ constexpr int wrmf_light_warp_doubles() { return kWrmfLightMax * KP + 2 * kWrmfLightMax * 17 + kWrmfLightMax + 2 * KP; }
template <int KP>
constexpr size_t wrmf_light_smem() {
    return sizeof(double) * (size_t)(KP * KP + kWrmfLightWarps * wrmf_light_warp_doubles<KP>());
}

template <int KP, bool LOSS>
__global__ void __launch_bounds__(kWrmfLightWarps * 32) wrmf_light_kernel(WrmfSide sd, const double* __restrict__ Binv) {
    constexpr int NCH = KP / 16;                         // chunks of 16 columns
    constexpr int DM = kWrmfLightMax;
    extern __shared__ __align__(16) double wrmf_smem[];
    double* Bs = wrmf_smem;                              // [KP][KP]  B^-1
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* Ys = Bs + KP * KP + warp * wrmf_light_warp_doubles<KP>();   // [16][KP]  rows of the other table, float64
    double* Zc = Ys + DM * KP;                           // [16][17]  one chunk of Z
    double* Ss = Zc + DM * 17;                           // [16][17]  S, eliminated in place
    double* rs = Ss + DM * 17;                           // [16]      rhs, then t
    double* qs = rs + DM;                                // [KP]      q = Y_u^T t
    double* xo = qs + KP;                                // [KP]      the row's solution before the update (loss)
    for (int e = threadIdx.x; e < KP * KP; e += blockDim.x) Bs[e] = Binv[e];
    __syncthreads();
    const unsigned full = 0xffffffffu;
    const int k = sd.k, ld = sd.ld;
    const int hh = lane >> 4, lc = lane & 15;            // (half, column) in the Z phase; (column half, row) in the M phase
    double loss = 0.0;
    const int64_t nw = (int64_t)gridDim.x * kWrmfLightWarps;
    for (int64_t row = sd.row_begin + (int64_t)blockIdx.x * kWrmfLightWarps + warp; row < sd.rows; row += nw) {
        const int64_t e0 = sd.indptr[row];
        const int d = (int)(sd.indptr[row + 1] - e0);
        if (d < 1 || d > DM) continue;                   // empty and longer rows belong to wrmf_solve_kernel
        __syncwarp();
        // ---- stage the d rows (float32 -> float64) and the weights ----
        int32_t my_idx = 0;
        double cw = 1.0;                                 // alpha r_ui of entry `lane`
        if (lane < d) { my_idx = sd.idx[e0 + lane]; cw = sd.alpha * (double)sd.cnt[e0 + lane]; }
        for (int e = 0; e < d; ++e) {
            const int64_t orow = __shfl_sync(full, my_idx, e);
            for (int c = lane; c < KP; c += 32) Ys[e * KP + c] = c < ld ? (double)__ldg(sd.other + orow * ld + c) : 0.0;
        }
        if (LOSS) for (int c = lane; c < KP; c += 32) xo[c] = c < k ? (double)sd.out[row * ld + c] : 0.0;
        if (lane < DM) rs[lane] = (1.0 + cw) / cw;
        __syncwarp();
        // ---- M = Y_u B^-1 Y_u^T, 16 columns of Z at a time ----
        double m[8];
#pragma unroll
        for (int f = 0; f < 8; ++f) m[f] = 0.0;
        double z[NCH][8];                                // Z[e = 2i + hh][16 ch + lc]
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
            for (int i = 0; i < 8; ++i) z[ch][i] = 0.0;
#pragma unroll 1
        for (int c = 0; c < KP; c += 2) {                // every y value is read once, every B^-1 value once per row
            double b0[NCH], b1[NCH];
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) { b0[ch] = Bs[c * KP + ch * 16 + lc]; b1[ch] = Bs[(c + 1) * KP + ch * 16 + lc]; }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (2 * i + hh < d) {
                    const double2 y = *reinterpret_cast<const double2*>(Ys + (2 * i + hh) * KP + c);
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) z[ch][i] = fma(y.x, b0[ch], fma(y.y, b1[ch], z[ch][i]));
                }
            }
        }
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
#pragma unroll
            for (int i = 0; i < 8; ++i) Zc[(2 * i + hh) * 17 + lc] = z[ch][i];
            __syncwarp();
            // lane (row lc, column half hh): m[ff] += sum_j Z[lc][j] Y[8 hh + ff][16 ch + j]
            if (lc < d) {
                for (int j = 0; j < 16; ++j) {
                    const double zj = Zc[lc * 17 + j];
#pragma unroll
                    for (int ff = 0; ff < 8; ++ff)
                        if (8 * hh + ff < d) m[ff] = fma(zj, Ys[(8 * hh + ff) * KP + ch * 16 + j], m[ff]);
                }
            }
            __syncwarp();
        }
        // ---- S = M + C^-1 in shared memory ----
        {
            const double cinv = 1.0 / __shfl_sync(full, cw, lc);          // 1 / (alpha r) of row lc
#pragma unroll
            for (int ff = 0; ff < 8; ++ff) Ss[lc * 17 + 8 * hh + ff] = m[ff] + ((8 * hh + ff == lc) ? cinv : 0.0);
        }
        __syncwarp();
        // ---- elimination (no pivoting: S is symmetric positive definite); lane (row lc, column half hh) ----
        for (int j = 0; j < d; ++j) {
            const double inv = wrmf_rcp(Ss[j * 17 + j]);
            const double l = Ss[lc * 17 + j] * inv;
            if (lc > j && lc < d) {
                const int f0 = hh ? (j + 1 > 8 ? j + 1 : 8) : j + 1, f1 = hh ? d : (d < 8 ? d : 8);
                for (int f = f0; f < f1; ++f) Ss[lc * 17 + f] = fma(-l, Ss[j * 17 + f], Ss[lc * 17 + f]);
                if (hh == 0) rs[lc] = fma(-l, rs[j], rs[lc]);
            }
            __syncwarp();
        }
        for (int j = d - 1; j >= 0; --j) {                // back substitution; rs[j] becomes t_j
            const double tj = rs[j] * wrmf_rcp(Ss[j * 17 + j]);
            __syncwarp();
            if (lane == j) rs[j] = tj;
            if (hh == 0 && lc < j) rs[lc] = fma(-Ss[lc * 17 + j], tj, rs[lc]);
            __syncwarp();
        }
        // ---- q = Y_u^T t (lane j: columns j, j + 32), x = B^-1 q ----
        double q[2] = {0.0, 0.0};
        for (int e = 0; e < d; ++e) {
            const double te = rs[e];
#pragma unroll
            for (int h = 0; h < 2; ++h) if (lane + 32 * h < KP) q[h] = fma(te, Ys[e * KP + lane + 32 * h], q[h]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) if (lane + 32 * h < KP) qs[lane + 32 * h] = q[h];
        __syncwarp();
        double x[2] = {0.0, 0.0};
        for (int c = 0; c < KP; ++c) {
            const double qc = qs[c];
#pragma unroll
            for (int h = 0; h < 2; ++h) if (lane + 32 * h < KP) x[h] = fma(qc, Bs[c * KP + lane + 32 * h], x[h]);
        }
        if (LOSS) {                                       // sum over the entries of (1 - x_old . y_e)^2, WRMF.py:49-50
            for (int e = 0; e < d; ++e) {
                double p = 0.0;
#pragma unroll
                for (int h = 0; h < 2; ++h) if (lane + 32 * h < KP) p = fma(xo[lane + 32 * h], Ys[e * KP + lane + 32 * h], p);
#pragma unroll
                for (int sft = 16; sft >= 1; sft >>= 1) p += __shfl_xor_sync(full, p, sft);
                if (lane == 0) { const double err = 1.0 - p; loss = fma(err, err, loss); }
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) if (lane + 32 * h < k) sd.out[row * ld + lane + 32 * h] = (float)x[h];
    }
    if (LOSS && lane == 0 && loss != 0.0) atomicAdd(sd.loss, loss);
}

// ---- pair counts and the track-major form of the play sets (WRMF.py:28-33, data/record.py:160-163) ----
__global__ void wrmf_count_kernel(const int64_t* __restrict__ ev_indptr, const int32_t* __restrict__ ev_items,
                                  const int64_t* __restrict__ uq_indptr, const int32_t* __restrict__ uq_items, int64_t m, int64_t T,
                                  const int32_t* __restrict__ hot_items, int32_t* __restrict__ cnt) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < T; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = m;                           // last u with ev_indptr[u] <= e
        while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (ev_indptr[mid] <= e) lo = mid; else hi = mid; }
        const int64_t u = lo;
        int32_t it = ev_items[e];
        if (it < 0) it = hot_items[-it - 1];              // hot positives are stored re-labelled -slot-1 (mark_hot_kernel)
        int64_t a = uq_indptr[u], b = uq_indptr[u + 1];
        while (a < b) { const int64_t mid = (a + b) >> 1; if (uq_items[mid] < it) a = mid + 1; else b = mid; }
        atomicAdd(cnt + a, 1);
    }
}
__global__ void wrmf_keys_kernel(const int64_t* __restrict__ uq_indptr, const int32_t* __restrict__ uq_items, int64_t m, int64_t nnz,
                                 uint64_t* __restrict__ keys) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = m;
        while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (uq_indptr[mid] <= p) lo = mid; else hi = mid; }
        keys[p] = ((uint64_t)(uint32_t)uq_items[p] << 32) | (uint32_t)lo;
    }
}

}  // namespace yue
