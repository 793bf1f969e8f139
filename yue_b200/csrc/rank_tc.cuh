// K3 (tcgen05 flavour): full-catalog scoring on the 5th-generation tensor cores with a fused,
// masked top-N epilogue, followed in the same kernel by the exact re-score (K5).
//
// Replaces the per-user body of evalRanking, base/IterativeRecommender.py:93-145, like
// rank_exact.cuh -- and returns the SAME ids and scores, bit for bit:
//
//   1. candidate pass.  A CTA owns 128 users (UMMA M = 128).  TMA streams Q in tiles of 128 tracks
//      (SWIZZLE_128B, fp32 in shared memory); one elected thread issues tcgen05.mma kind::tf32
//      (the tensor core reads the fp32 bits and drops the low mantissa bits: no conversion pass)
//      into a double-buffered TMEM accumulator; four epilogue warps read it back with tcgen05.ld
//      -- one thread per user row, 32 columns per load -- reduce each chunk to its maximum and
//      compare with the row's running threshold.  After the first tiles > 99 % of chunks are
//      rejected by that one compare.  Survivors are checked against the user's sorted play row with
//      a monotone cursor (tiles arrive in track order) and pushed into a per-row buffer in shared
//      memory; a full buffer is compacted by the whole warp.
//   2. the tf32 scores are approximate, so the threshold is tau - 2*eps with tau the N-th best
//      approximate score so far and eps = c * |p_u| * max_t |q_t| a bound on the tf32 error
//      (c = 2^-9 for two truncated operands + accumulation slack).  Every track whose EXACT score is
//      among the N best then survives: exact(x) >= sigma  =>  approx(x) >= sigma - eps >= tau - 2 eps
//      (sigma = N-th best exact score >= tau - eps because N tracks have approx >= tau).
//   3. exact pass.  The surviving <= 64 candidates of a row are re-scored with the canonical fp32
//      FMA chain and sorted by (score desc, id asc) -- identical to the exact kernel.
//   4. a row whose buffer cannot hold its candidates (too many near-ties) is reported in fail_rows
//      and the caller re-runs it through rank_exact_kernel.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
// warps 2..5 = epilogue (TMEM lane quarter = warp % 4).
#pragma once
#include <cuda.h>            // CUtensorMap + enums only; the encoder is fetched at run time
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "rank_exact.cuh"

namespace yue {

constexpr int kTcThreads = 192;
constexpr int kTcBM = 128;            // users per CTA (UMMA M)
// tracks per tile (UMMA N) and TMA -> MMA shared-memory stages are template parameters of the kernel:
//   d <= 64:  BN = 128, 3 stages of 32 KB;   64 < d <= 128:  BN = 64, 2 stages of 32 KB (A alone is 64 KB)
// candidate slots per row (template parameter CAP of the kernel): CAP - 32 kept + 32 new per chunk.  64 for N <= 12
// (fastest: 10.8 ms per wave at N = 10); 96 for larger N (at N = 20 a row keeps ~27-45 near-ties: with 64 slots
// nearly every row spilled and fell back to the exact kernel, 202 ms instead of 12.8 ms).
constexpr int kTcAcc = 4;             // TMEM accumulator stages (4 x 128 columns = all of TMEM)
constexpr int kTcOvfCap = 1024;       // candidate slots of a row that overflowed into the global pool
constexpr int kTcOvfRowsMin = 512;    // rows the spill pool can take per launch: at least this, else B / 16
constexpr int kTcBoxBytes = 128 * 128;   // one TMA box: 128 rows x 128 B
constexpr float kTf32ErrCoef = 2.1e-3f;  // 2^-9 (two operands truncated to 10 mantissa bits) + slack

struct RankTcParams {
    int kblocks;                // 128-byte k-blocks per row: ceil(ld / 32), 1..4
    int ntiles;
    int n_items;
    int64_t B;
    int N;
    int ld, d;
    const float* Psel;          // [Bpad, ld] gathered user rows
    const float* Q;
    const float* pnorm;         // [Bpad] |p_u|
    const float* qmax;          // max_t |q_t|
    const int32_t* users;       // [B]
    const int64_t* uq_indptr;
    const int32_t* uq_items;
    int32_t* ids_out;
    float* scores_out;
    int* fail_count;
    int32_t* fail_rows;
    uint64_t* ovf_pool;         // [ovf_rows, kTcOvfCap] spill space for rows with too many near-ties
    int* ovf_next;
    int ovf_rows;
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// the same box delivered to the same shared-memory offset of every CTA in `mask` (and its bytes signalled on the barrier at
// the same offset in each of them): one read of L2 serves the whole cluster
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// arrive on the barrier at this offset in every CTA of `mask` once the MMAs issued so far have retired
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" :: "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}" :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// true on exactly one lane of a converged warp (the MMA warp runs converged so that ptxas keeps the MMA's operands in
// uniform registers; issuing from `if (lane == 0)` costs a uniform-register waterfall loop per MMA)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n .reg .pred P;\n elect.sync _|P, 0xffffffff;\n selp.b32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}
// One 128-byte k-block = 4 MMAs (K = 8 tf32 each), issued from ONE asm block with descriptors the caller precomputed: the
// issuing thread is a single warp executing dependent scalar code at ~6 cycles per instruction, and round 1's loop
// (descriptor built per MMA, tile counters by division, one uniform-register waterfall per MMA: ~170 instructions per tile)
// took ~1000 cycles per tile against 536 of tensor-pipe work (profiles/ncu_rank_r2.md).  Along the swizzle row the k-th
// MMA's operands start 32 bytes further: +2 in the descriptor's 16-byte address field.
__device__ __forceinline__ void tc_mma_tf32_x4(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc_first) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 a1, a2, a3, b1, b2, b3;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "add.u64 a1, %1, 2;\n add.u64 a2, %1, 4;\n add.u64 a3, %1, 6;\n"
        "add.u64 b1, %2, 2;\n add.u64 b2, %2, 4;\n add.u64 b3, %2, 6;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], a1, b1, %3, 1;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], a2, b2, %3, 1;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], a3, b3, %3, 1;\n"
        "}" :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc_first) : "memory");
}
// The same with A read from TENSOR MEMORY (lane = row of A, one 32-bit column per tf32 element): the user block is the same
// for every tile, and with both operands in shared memory the K = 8 tf32 MMA reads 8 KB per 65 cycles -- all of the SM's
// shared-memory bandwidth, on top of which the TMA writes the next Q tile (32 KB): the bare pipeline ran at 750 cycles per
// tile = (8 x 8 KB + 32 KB) / 128 B/clk.  A in TMEM halves the operand traffic.
__device__ __forceinline__ void tc_mma_tf32_ts_x4(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc_first) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 b1, b2, b3;\n"
        ".reg .b32 a1, a2, a3;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "add.u32 a1, %1, 8;\n add.u32 a2, %1, 16;\n add.u32 a3, %1, 24;\n"
        "add.u64 b1, %2, 2;\n add.u64 b2, %2, 4;\n add.u64 b3, %2, 6;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [a1], b1, %3, 1;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [a2], b2, %3, 1;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [a3], b3, %3, 1;\n"
        "}" :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc_first) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns written from registers: thread t of the warp writes TMEM lane (base lane + t)
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
           "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
           "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets TMEM lane (base lane + t)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// issue only (no wait): 64 consecutive columns
__device__ __forceinline__ void tc_ld64_issue(uint32_t taddr, uint32_t (&r)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
          "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
          "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
          "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
          "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major operand, SWIZZLE_128B: rows of 128 B, 8-row atoms 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);      // start address, 16-byte units   [0,14)
    d |= (uint64_t)1 << 16;                                 // leading byte offset (unused here) [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                       // stride byte offset: 8 rows x 128 B [32,46)
    d |= (uint64_t)1 << 46;                                 // descriptor version 1 (sm_100)     [46,48)
    d |= (uint64_t)2 << 61;                                 // layout type SWIZZLE_128B          [61,64)
    return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 128 (cute::UMMA::InstrDescriptor)
template <int BN> __host__ __device__ constexpr uint32_t idesc_tf32() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
}

// ---- the rare path of the epilogue, kept OUT of line ------------------------------------------
// ncu on the first version: 73 % of the epilogue's stall samples were `no_inst` -- the 32-way
// unrolled push code, inlined at four sites, made a 218 KB kernel that thrashed the instruction
// cache.  Now an unrolled slot is a compare and a predicated call.
struct TcRow {                 // per-thread (= per user row) state the rare path works on
    int64_t mcur, mend;        // cursor into the user's sorted play row (tiles arrive in track order)
    int mval;                  // uq_items[mcur] or INT32_MAX
    int cnt;                   // entries in the row's shared-memory buffer
    int ovf_slot, ovf_cnt;     // >= 0: the row spills into the global pool
    int fail;                  // row must be redone by the exact kernel
};
// push one surviving score; returns true when the row has to give up (pool row full)
__device__ __noinline__ bool tc_push(TcRow* st, uint64_t* K, const int32_t* __restrict__ uq_items,
                                     uint64_t* __restrict__ ovf_pool, float sc, int id) {
    while (st->mval < id) { ++st->mcur; st->mval = st->mcur < st->mend ? uq_items[st->mcur] : INT32_MAX; }
    if (st->mval == id) return false;                               // a training track of this user
    if (st->ovf_slot < 0) { K[st->cnt++] = make_key(sc, id); return false; }   // cnt <= 32 before the chunk
    if (st->ovf_cnt < kTcOvfCap) { ovf_pool[(size_t)st->ovf_slot * kTcOvfCap + st->ovf_cnt++] = make_key(sc, id); return false; }
    st->fail = 1;
    return true;
}
// whole warp: compact every row of this warp whose buffer is more than half full; returns the
// (possibly raised) push threshold of the calling lane's row
template <int CAP>
__device__ __noinline__ float tc_compact(TcRow* st, uint64_t* keys_quarter, int lane, int N, float eps2, float thr,
                                         uint64_t* __restrict__ ovf_pool, int* __restrict__ ovf_next, int ovf_rows) {
    unsigned full;
    constexpr int kTcPer = CAP / 32;
    while ((full = __ballot_sync(0xffffffffu, st->cnt > CAP - 32)) != 0u) {
        const int src = __ffs(full) - 1;
        const int cc = __shfl_sync(0xffffffffu, st->cnt, src);
        const float e2 = __shfl_sync(0xffffffffu, eps2, src);
        uint64_t* R = keys_quarter + (size_t)src * CAP;
        uint64_t mine[kTcPer];
        int rk[kTcPer];
#pragma unroll
        for (int q = 0; q < kTcPer; ++q) { mine[q] = lane + 32 * q < cc ? R[lane + 32 * q] : ~0ull; rk[q] = 0; }
        for (int e = 0; e < cc; ++e) {
            const uint64_t k = R[e];
#pragma unroll
            for (int q = 0; q < kTcPer; ++q) rk[q] += k < mine[q];
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < kTcPer; ++q) if (lane + 32 * q < cc) R[rk[q]] = mine[q];
        __syncwarp();
        const float lim = key_score(R[N - 1]) - e2;                 // cc > CAP - 32 >= N
        int kept = 0;
#pragma unroll
        for (int q = 0; q < kTcPer; ++q)
            kept += __popc(__ballot_sync(0xffffffffu, lane + 32 * q < cc && key_score(R[lane + 32 * q]) >= lim));
        int slot = -1;
        if (kept > CAP - 32) {                 // too many near-ties for shared memory: spill the row
            if (lane == 0) slot = atomicAdd(ovf_next, 1);
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (slot < ovf_rows) {
                uint64_t* G = ovf_pool + (size_t)slot * kTcOvfCap;
#pragma unroll
                for (int q = 0; q < kTcPer; ++q) if (lane + 32 * q < kept) G[lane + 32 * q] = R[lane + 32 * q];
            }
        }
        if (lane == src) {
            thr = lim;
            if (kept <= CAP - 32) st->cnt = kept;
            else if (slot < ovf_rows) { st->ovf_slot = slot; st->ovf_cnt = kept; st->cnt = 0; }   // threshold frozen from here on
            else { st->fail = 1; thr = INFINITY; st->cnt = 0; }
        }
        __syncwarp();
    }
    return thr;
}

// EXP: timing experiments only (YUE_RANK_EXP; results are NOT valid for EXP = 1): 1 = the epilogue reads TMEM and frees the
// stage but looks at nothing (what the TMA -> MMA -> TMEM-read pipeline does on its own); 2 = both halves of a tile are
// requested before the first wait.
// CL: CTAs per cluster (1, 2 or 4).  The CTAs of a cluster walk the catalog together: each loads 1/CL of every Q tile and
// TMA multicasts it to all of them, so a tile is read from L2 once per cluster instead of once per CTA (every CTA streams
// all of Q: 11.8 TB/s over a wave with CL = 1, the bound of the bare pipeline -- profiles/ncu_rank_r2.md).  A stage may be
// refilled only when the MMAs of ALL the cluster's CTAs have read it: their commits arrive on every CTA's `empty` barrier.
// ATM: the user block (A) lives in tensor memory instead of shared memory (d <= 64: 3 accumulator stages of 128 columns + 64
// columns of A; see tc_mma_tf32_ts_x4).
template <int CAP, int BN, int kTcStages, int EXP = 0, int CL = 1, bool ATM = false>
__global__ void __launch_bounds__(kTcThreads, 1)
rank_tc_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmQ, const RankTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_tc_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_tc_raw) + 1023) & ~(uintptr_t)1023);
    const int KB = p.kblocks;
    constexpr int NACC = ATM ? 3 : kTcAcc;                   // TMEM accumulator stages
    constexpr uint32_t kAtmCol = NACC * BN;                  // first TMEM column of A (ATM)
    constexpr uint32_t kBBox = BN * 128;                     // one TMA box of Q: BN rows x 128 B
    const uint32_t a_bytes = ATM ? 0u : (uint32_t)KB * kTcBoxBytes, stage_bytes = (uint32_t)KB * kBBox;
    uint8_t* sA = smem;
    uint8_t* sB = sA + a_bytes;
    constexpr int kTcPer = CAP / 32;
    uint64_t* keys = reinterpret_cast<uint64_t*>(sB + kTcStages * stage_bytes);        // [128][CAP]
    uint64_t* bars = keys + kTcBM * CAP;
    // barriers: 0 = A landed, 1..S = B full, S+1..2S = B empty, then kTcAcc TMEM full, kTcAcc TMEM empty
    const uint32_t bar_a = smem_u32(bars);
    auto bar_full = [&](int s) { return smem_u32(bars + 1 + s); };
    auto bar_empty = [&](int s) { return smem_u32(bars + 1 + kTcStages + s); };
    auto bar_tfull = [&](int t) { return smem_u32(bars + 1 + 2 * kTcStages + t); };
    auto bar_tempty = [&](int t) { return smem_u32(bars + 1 + 2 * kTcStages + kTcAcc + t); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 2 * kTcStages + 2 * kTcAcc);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    constexpr uint16_t kClusterMask = (uint16_t)((1u << CL) - 1u);
    if (threadIdx.x == 0) {
        mbar_init(bar_a, ATM ? 128 : 1);
        for (int s = 0; s < kTcStages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), CL); }
        for (int t = 0; t < NACC; ++t) { mbar_init(bar_tfull(t), 1); mbar_init(bar_tempty(t), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmP) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmQ) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (CL > 1) cluster_sync_all();                     // every CTA's barriers exist before a peer signals them
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            if (!ATM) {
                mbar_expect_tx(bar_a, a_bytes);
                for (int kb = 0; kb < KB; ++kb)
                    tma_load_2d(smem_u32(sA + kb * kTcBoxBytes), &tmP, bar_a, kb * 32, blockIdx.x * kTcBM);
            }
            for (int j = 0; j < p.ntiles; ++j) {
                const int s = j % kTcStages;
                const uint32_t ph = (uint32_t)(j / kTcStages) & 1u;
                mbar_wait(bar_empty(s), ph ^ 1u);
                mbar_expect_tx(bar_full(s), stage_bytes);       // the whole stage: this CTA's part and its peers'
                for (int kb = 0; kb < KB; ++kb) {
                    if (CL == 1) tma_load_2d(smem_u32(sB + s * stage_bytes + kb * kBBox), &tmQ, bar_full(s), kb * 32, j * BN);
                    else tma_load_2d_mc(smem_u32(sB + s * stage_bytes + kb * kBBox) + crank * (uint32_t)(BN / CL * 128), &tmQ, bar_full(s),
                                        kb * 32, j * BN + (int)crank * (BN / CL), kClusterMask);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the warp stays converged, one elected lane issues =====
        {
            const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
            mbar_wait(bar_a, 0);
            // everything that does not depend on the tile is computed once: operand descriptors (A per k-block, B of stage 0;
            // a stage / k-block further is a constant in the 16-byte address field); stage and accumulator indices, phases
            // and the stage offset advance by increments
            const uint64_t adesc0 = umma_desc_sw128(smem_u32(sA));
            const uint64_t bdesc_base = umma_desc_sw128(smem_u32(sB));
            const uint32_t stage16 = stage_bytes >> 4, kb16 = kBBox >> 4, ka16 = kTcBoxBytes >> 4;
            int s = 0, t = 0;
            uint32_t ph = 0, tph = 0, boff = 0;
            for (int j = 0; j < p.ntiles; ++j) {
                mbar_wait(bar_tempty(t), tph ^ 1u);
                mbar_wait(bar_full(s), ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t dcol = tbase + (uint32_t)t * BN;
                    const uint64_t bd = bdesc_base + boff;      // stays inside the 14-bit address field (smem < 256 KB)
                    if (ATM) {
                        tc_mma_tf32_ts_x4(dcol, tbase + kAtmCol, bd, idesc_tf32<BN>(), 0u);
                        for (int kb = 1; kb < KB; ++kb)
                            tc_mma_tf32_ts_x4(dcol, tbase + kAtmCol + (uint32_t)kb * 32u, bd + (uint32_t)kb * kb16, idesc_tf32<BN>(), 1u);
                    } else {
                        tc_mma_tf32_x4(dcol, adesc0, bd, idesc_tf32<BN>(), 0u);
                        for (int kb = 1; kb < KB; ++kb)         // up to 4 k-blocks of 32 floats (num.factors <= 128)
                            tc_mma_tf32_x4(dcol, adesc0 + (uint32_t)kb * ka16, bd + (uint32_t)kb * kb16, idesc_tf32<BN>(), 1u);
                    }
                    if (CL == 1) tc_commit(bar_empty(s));   // smem stage free once these MMAs retire
                    else tc_commit_mc(bar_empty(s), kClusterMask);
                    tc_commit(bar_tfull(t));            // accumulator ready for the epilogue
                }
                __syncwarp();
                boff += stage16;
                if (++s == kTcStages) { s = 0; ph ^= 1u; boff = 0; }
                if (++t == NACC) { t = 0; tph ^= 1u; }
            }
        }
    } else {
        // ===== epilogue: one thread per user row =====
        const int quarter = warp & 3;                   // TMEM lanes this warp may read
        const int r = quarter * 32 + lane;
        const int64_t b = (int64_t)blockIdx.x * kTcBM + r;
        const bool valid = b < p.B;
        uint64_t* K = keys + (size_t)r * CAP;
        const float eps2 = valid ? 2.f * kTf32ErrCoef * p.pnorm[b] * (*p.qmax) : 0.f;
        float thr = valid ? -INFINITY : INFINITY;       // push threshold = tau - 2 eps
        TcRow st;
        st.mcur = 0; st.mend = 0; st.cnt = 0; st.ovf_slot = -1; st.ovf_cnt = 0; st.fail = 0;
        if (valid) { const int u = p.users[b]; st.mcur = p.uq_indptr[u]; st.mend = p.uq_indptr[u + 1]; }
        st.mval = st.mcur < st.mend ? p.uq_items[st.mcur] : INT32_MAX;
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
        uint64_t* keys_quarter = keys + (size_t)(quarter * 32) * CAP;

        // examine 32 scores of columns [col0, col0+32): push survivors, then compact crowded rows.  Entered by the whole warp
        // when ANY of its 32 rows has a chunk maximum above its threshold (~3 % of the chunks once the thresholds have
        // settled), and a warp that is late with a tile holds up the MMA of that TMEM stage for all four quarters -- so the
        // path is kept short: a row first looks at the maxima of its four groups of 8 scores and scans only a group that can
        // hold a survivor (usually one row, one group), and the compaction is called only when some row's buffer is crowded.
#define TC_RARE(v, col0, m)                                                                          \
        do {                                                                                         \
            if ((m) >= thr) {                                                                        \
                _Pragma("unroll")                                                                    \
                for (int g = 0; g < 4; ++g) {                                                        \
                    float mg = __uint_as_float((v)[8 * g]);                                          \
                    _Pragma("unroll")                                                                \
                    for (int x = 1; x < 8; ++x) mg = fmaxf(mg, __uint_as_float((v)[8 * g + x]));     \
                    if (mg >= thr) {                                                                 \
                        _Pragma("unroll")                                                            \
                        for (int x = 8 * g; x < 8 * g + 8; ++x) {                                    \
                            const float sc = __uint_as_float((v)[x]);                                \
                            if (sc >= thr && (col0) + x < p.n_items)                                 \
                                if (tc_push(&st, K, p.uq_items, p.ovf_pool, sc, (col0) + x)) thr = INFINITY; \
                        }                                                                            \
                    }                                                                                \
                }                                                                                    \
            }                                                                                        \
            __syncwarp();                                                                            \
            if (__any_sync(0xffffffffu, st.cnt > CAP - 32))                                          \
                thr = tc_compact<CAP>(&st, keys_quarter, lane, p.N, eps2, thr, p.ovf_pool, p.ovf_next, p.ovf_rows); \
        } while (0)

        if (ATM) {
            // this thread's row of the user block -> TMEM lane r, columns kAtmCol ... (fp32 bits; the MMA reads them as tf32)
            const float* prow = p.Psel + (size_t)((int64_t)blockIdx.x * kTcBM + r) * p.ld;        // rows past B are zero rows
            for (int kb = 0; kb < KB; ++kb) {
                uint32_t a[32];
#pragma unroll
                for (int x = 0; x < 32; x += 4) {
                    const bool in = kb * 32 + x < p.ld;
                    const float4 v = in ? __ldg(reinterpret_cast<const float4*>(prow + kb * 32 + x)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    a[x] = __float_as_uint(v.x); a[x + 1] = __float_as_uint(v.y); a[x + 2] = __float_as_uint(v.z); a[x + 3] = __float_as_uint(v.w);
                }
                tc_st32(trow + kAtmCol + (uint32_t)kb * 32u, a);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(bar_a);
        }
        int t = 0;
        uint32_t tph = 0;
        for (int j = 0; j < p.ntiles; ++j, t = (t + 1 == NACC ? 0 : t + 1), tph ^= (t == 0 ? 1u : 0u)) {
            const int i0 = j * BN;
            mbar_wait(bar_tfull(t), tph);
            tc_fence_after();
            uint32_t va[64], vb[64];
            tc_ld64_issue(trow + (uint32_t)(t * BN), va);
            if (EXP != 2) tc_ld_wait();
            if (BN == 128) tc_ld64_issue(trow + (uint32_t)(t * BN + 64), vb);   // in flight while the first half is reduced
            if (EXP == 2) tc_ld_wait();
            if (EXP == 1) {
                tc_ld_wait();
                tc_fence_before();
                mbar_arrive(bar_tempty(t));
                if (va[lane] == 0x7fc00001u && vb[lane] == 0x7fc00002u) thr = 0.f;      // keep the loads alive
                continue;
            }
            float m0 = __uint_as_float(va[0]), m1 = __uint_as_float(va[32]);
#pragma unroll
            for (int x = 1; x < 32; ++x) { m0 = fmaxf(m0, __uint_as_float(va[x])); m1 = fmaxf(m1, __uint_as_float(va[32 + x])); }
            if (__any_sync(0xffffffffu, m0 >= thr)) TC_RARE(va, i0, m0);
            if (__any_sync(0xffffffffu, m1 >= thr)) TC_RARE(va + 32, i0 + 32, m1);
            tc_ld_wait();
            tc_fence_before();
            mbar_arrive(bar_tempty(t));                                 // TMEM stage free: the rest works on registers
            if (BN != 128) continue;
            float m2 = __uint_as_float(vb[0]), m3 = __uint_as_float(vb[32]);
#pragma unroll
            for (int x = 1; x < 32; ++x) { m2 = fmaxf(m2, __uint_as_float(vb[x])); m3 = fmaxf(m3, __uint_as_float(vb[32 + x])); }
            if (__any_sync(0xffffffffu, m2 >= thr)) TC_RARE(vb, i0 + 64, m2);
            if (__any_sync(0xffffffffu, m3 >= thr)) TC_RARE(vb + 32, i0 + 96, m3);
        }
#undef TC_RARE
        const int cnt = st.cnt, ovf_slot = st.ovf_slot, ovf_cnt = st.ovf_cnt;
        const bool fail = st.fail != 0;

        // ---- exact pass: re-score the survivors with the fp32 FMA chain, sort, write -------------
        __syncwarp();
        for (int rr = 0; rr < 32; ++rr) {
            const int64_t bb = (int64_t)blockIdx.x * kTcBM + quarter * 32 + rr;
            if (bb >= p.B) break;
            int cc = __shfl_sync(0xffffffffu, cnt, rr);
            const bool ff = __shfl_sync(0xffffffffu, (int)fail, rr) != 0;
            const int oslot = __shfl_sync(0xffffffffu, ovf_slot, rr), ocnt = __shfl_sync(0xffffffffu, ovf_cnt, rr);
            if (ff) {
                if (lane == 0) p.fail_rows[atomicAdd(p.fail_count, 1)] = (int32_t)bb;
                continue;
            }
            uint64_t* R = keys + (size_t)(quarter * 32 + rr) * CAP;
            const float* pu = p.Psel + (size_t)bb * p.ld;
            if (oslot >= 0) {
                // spilled row: stream the pool entries through the 64-slot buffer, 32 at a time, keeping
                // the N best EXACT scores in R[0..N)
                const uint64_t* G = p.ovf_pool + (size_t)oslot * kTcOvfCap;
                int have = 0;
                for (int base = 0; base < ocnt; base += 32) {
                    const int e = base + lane;
                    uint64_t nk = ~0ull;
                    if (e < ocnt) {
                        const int id = key_id(G[e]);
                        nk = make_key(score_fma32(pu, p.Q + (size_t)id * p.ld, p.d), id);
                    }
                    const int take = min(32, ocnt - base);
                    __syncwarp();
                    if (lane < take) R[have + lane] = nk;
                    __syncwarp();
                    have = compact_row<CAP>(R, have + take, p.N, lane);
                    __syncwarp();
                }
                for (int x = lane; x < p.N; x += 32) {
                    const bool ok = x < have;
                    const uint64_t k = ok ? R[x] : 0ull;
                    p.ids_out[bb * p.N + x] = ok ? key_id(k) : -1;
                    p.scores_out[bb * p.N + x] = ok ? key_score(k) : -INFINITY;
                }
                continue;
            }
            uint64_t nk[kTcPer];
#pragma unroll
            for (int h = 0; h < kTcPer; ++h) {
                const int e = lane + 32 * h;
                nk[h] = 0;
                if (e < cc) {
                    const int id = key_id(R[e]);
                    nk[h] = make_key(score_fma32(pu, p.Q + (size_t)id * p.ld, p.d), id);
                }
            }
            __syncwarp();
#pragma unroll
            for (int h = 0; h < kTcPer; ++h) if (lane + 32 * h < cc) R[lane + 32 * h] = nk[h];
            __syncwarp();
            const int nc = compact_row<CAP>(R, cc, p.N, lane);
            for (int x = lane; x < p.N; x += 32) {
                const bool ok = x < nc;
                const uint64_t k = ok ? R[x] : 0ull;
                p.ids_out[bb * p.N + x] = ok ? key_id(k) : -1;
                p.scores_out[bb * p.N + x] = ok ? key_score(k) : -INFINITY;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                     // no CTA leaves while a peer may still signal its barriers
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- helpers of the TC path ---------------------------------------------------------------------
__global__ void tc_gather_rows_kernel(const float* __restrict__ P, const int32_t* __restrict__ users, int64_t B,
                                      int64_t Bpad, int ld, float* __restrict__ Psel, float* __restrict__ pnorm) {
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= Bpad) return;
    float acc = 0.f;
    for (int k = lane; k < ld; k += 32) {
        const float v = row < B ? P[(size_t)users[row] * ld + k] : 0.f;
        Psel[(size_t)row * ld + k] = v;
        acc += v * v;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if (lane == 0) pnorm[row] = sqrtf(acc) * 1.0001f;
}
__global__ void tc_row_norm_max_kernel(const float* __restrict__ Q, int64_t n, int ld, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    float best = 0.f;
    for (int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; row < n;
         row += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        float acc = 0.f;
        for (int k = lane; k < ld; k += 32) { const float v = Q[(size_t)row * ld + k]; acc += v * v; }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
        best = fmaxf(best, acc);
    }
    if (lane == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(sqrtf(best) * 1.0001f));   // >= 0: int order = float order
}
__global__ void tc_scatter_rows_kernel(const int32_t* __restrict__ rows, int nrows, int N, const int32_t* __restrict__ ids,
                                       const float* __restrict__ sc, int32_t* __restrict__ ids_out, float* __restrict__ sc_out) {
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < nrows * N; x += gridDim.x * blockDim.x) {
        const int64_t dst = (int64_t)rows[x / N] * N + x % N;
        ids_out[dst] = ids[x];
        sc_out[dst] = sc[x];
    }
}
__global__ void tc_gather_users_kernel(const int32_t* __restrict__ users, const int32_t* __restrict__ rows, int nrows,
                                       int32_t* __restrict__ out) {
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < nrows; x += gridDim.x * blockDim.x) out[x] = users[rows[x]];
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct RankTcState {
    bool q_dirty = true;
    PFN_tmapEncodeTiled encode = nullptr;
    float* qmax = nullptr;
    float* psel = nullptr; size_t psel_cap = 0;
    float* pnorm = nullptr; size_t pnorm_cap = 0;
    int* fail_count = nullptr;
    int32_t* fail_rows = nullptr; size_t fail_cap = 0;
    int32_t* fb_users = nullptr; int32_t* fb_ids = nullptr; float* fb_scores = nullptr; size_t fb_cap = 0;
    uint64_t* ovf_pool = nullptr;
    int* ovf_next = nullptr;
    int64_t last_fail = 0;       // rows sent to the exact kernel by the last call (diagnostics)
    int64_t last_spill = 0;      // rows that spilled into the pool during the last call
    size_t ovf_rows = 0;
};

// a row keeps its N best plus the near-ties (within 2 eps of the N-th) in kTcCap - 32 = 64 slots
inline bool rank_tc_supported(int k, int N) { return k >= 1 && k <= 128 && N >= 1 && N <= 32; }

inline void rank_tc_release(RankTcState& st) {
    for (void* p : {(void*)st.qmax, (void*)st.psel, (void*)st.pnorm, (void*)st.fail_count, (void*)st.fail_rows,
                    (void*)st.fb_users, (void*)st.fb_ids, (void*)st.fb_scores, (void*)st.ovf_pool, (void*)st.ovf_next})
        if (p) cudaFree(p);
    st = RankTcState();
}

template <class T>
static cudaError_t tc_grow(T*& p, size_t& cap, size_t need) {
    if (need <= cap && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cudaError_t e = cudaMalloc((void**)&p, need * sizeof(T));
    cap = e == cudaSuccess ? need : 0;
    return e;
}

static bool tc_make_map(PFN_tmapEncodeTiled enc, CUtensorMap* map, const float* base, uint64_t rows, int ld, unsigned box_rows = 128u) {
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {32u, box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Runs the TC path on device buffers.  `fallback(d_users, B, d_ids, d_scores)` must run the exact
// kernel for the given rows.  Returns 0 or a YUE_E_* code (2 = CUDA).
template <class Fallback>
inline int rank_tc_run(RankTcState& st, cudaStream_t stream, int sm_count, const float* P, const float* Q, int ld, int d,
                       int n_items, const int32_t* d_users, int64_t B, int N, const int64_t* uq_indptr,
                       const int32_t* uq_items, int32_t* d_ids, float* d_scores, std::string& err, int64_t& launches,
                       Fallback fallback) {
#define TC_CK(call)                                                                 \
    do {                                                                            \
        cudaError_t e_ = (call);                                                    \
        if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); return 2; } \
    } while (0)
    if (!st.encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        TC_CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) { err = "cuTensorMapEncodeTiled not available"; return 6; }
        st.encode = (PFN_tmapEncodeTiled)fn;
        TC_CK(cudaMalloc((void**)&st.qmax, sizeof(float)));
        TC_CK(cudaMalloc((void**)&st.fail_count, sizeof(int)));
        TC_CK(cudaMalloc((void**)&st.ovf_next, sizeof(int)));
    }
    // CTAs per cluster sharing each Q tile by TMA multicast (YUE_RANK_CLUSTER = 1, 2, 4; see rank_tc_kernel)
    int cl = 1;                                           // measured on B200: 2 = no gain (L2 is not the bound), 4 = half the clusters fit
    if (const char* e = getenv("YUE_RANK_CLUSTER")) cl = atoi(e);
    if (cl != 2 && cl != 4) cl = 1;
    if (B <= kTcBM) cl = 1;                               // a single user block: nobody to share with
    const int64_t Bpad = (B + kTcBM * cl - 1) / (kTcBM * cl) * (kTcBM * cl);   // whole clusters; rows past B are zero rows
    TC_CK(tc_grow(st.ovf_pool, st.ovf_rows, (size_t)std::max<int64_t>(kTcOvfRowsMin, B / 16) * kTcOvfCap));
    size_t cap_tmp = st.psel_cap;
    TC_CK(tc_grow(st.psel, st.psel_cap, (size_t)Bpad * ld));
    (void)cap_tmp;
    TC_CK(tc_grow(st.pnorm, st.pnorm_cap, (size_t)Bpad));
    TC_CK(tc_grow(st.fail_rows, st.fail_cap, (size_t)Bpad));
    if (st.q_dirty) {
        TC_CK(cudaMemsetAsync(st.qmax, 0, sizeof(float), stream));
        tc_row_norm_max_kernel<<<sm_count * 8, 256, 0, stream>>>(Q, n_items, ld, st.qmax);
        ++launches;
        st.q_dirty = false;
    }
    TC_CK(cudaMemsetAsync(st.fail_count, 0, sizeof(int), stream));
    TC_CK(cudaMemsetAsync(st.ovf_next, 0, sizeof(int), stream));
    tc_gather_rows_kernel<<<(unsigned)((Bpad * 32 + 255) / 256), 256, 0, stream>>>(P, d_users, B, Bpad, ld, st.psel, st.pnorm);
    ++launches;
    TC_CK(cudaGetLastError());

    CUtensorMap tmP, tmQ;
    // A in tensor memory (default): the user block costs no shared memory, so every width takes 128-track tiles (3 stages of
    // 32 KB for d <= 64, 2 stages of up to 64 KB above); YUE_RANK_ATM=0 / clusters: round 1's layout (A in shared memory)
    const bool atm = cl == 1 && !(getenv("YUE_RANK_ATM") && atoi(getenv("YUE_RANK_ATM")) == 0);
    const int bn = (ld <= 64 || atm) ? 128 : 64;
    int stages = ld <= 64 ? 3 : 2;
    if (!tc_make_map(st.encode, &tmP, st.psel, (uint64_t)Bpad, ld) || !tc_make_map(st.encode, &tmQ, Q, (uint64_t)n_items, ld, (unsigned)(bn / cl))) {
        err = "cuTensorMapEncodeTiled failed";
        return 2;
    }
    RankTcParams p{};
    p.kblocks = (ld + 31) / 32;
    p.ntiles = (n_items + bn - 1) / bn;
    p.n_items = n_items; p.B = B; p.N = N; p.ld = ld; p.d = d;
    p.Psel = st.psel; p.Q = Q; p.pnorm = st.pnorm; p.qmax = st.qmax; p.users = d_users;
    p.uq_indptr = uq_indptr; p.uq_items = uq_items; p.ids_out = d_ids; p.scores_out = d_scores;
    p.fail_count = st.fail_count; p.fail_rows = st.fail_rows;
    p.ovf_pool = st.ovf_pool; p.ovf_next = st.ovf_next; p.ovf_rows = (int)(st.ovf_rows / kTcOvfCap);
    const int cap = N <= 12 ? 64 : 96;
    if (bn == 128 && cap == 64 && getenv("YUE_RANK_STAGES4")) stages = 4;       // experiment: a fourth Q stage (measured: no gain)
    const size_t smem = 1024 + (size_t)p.kblocks * ((atm ? 0 : kTcBoxBytes) + (size_t)stages * bn * 128) + (size_t)kTcBM * cap * 8 + 256;
    const unsigned grid = (unsigned)(Bpad / kTcBM);
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(grid, 1, 1); lc.blockDim = dim3(kTcThreads, 1, 1); lc.dynamicSmemBytes = smem; lc.stream = stream;
    cudaLaunchAttribute lattr[1];
    lattr[0].id = cudaLaunchAttributeClusterDimension;
    lattr[0].val.clusterDim.x = (unsigned)cl; lattr[0].val.clusterDim.y = 1; lattr[0].val.clusterDim.z = 1;
    lc.attrs = lattr; lc.numAttrs = 1;
#define TC_LAUNCH_K(K_)                                                                                                    \
    do {                                                                                                                   \
        TC_CK(cudaFuncSetAttribute(K_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                           \
        TC_CK(cudaLaunchKernelEx(&lc, K_, tmP, tmQ, p));                                                                   \
    } while (0)
#define TC_LAUNCH(CAP_, BN_, ST_)                                                                                          \
    do {                                                                                                                   \
        if (cl == 2) TC_LAUNCH_K((rank_tc_kernel<CAP_, BN_, ST_, 0, 2>));                                                  \
        else if (cl == 4) TC_LAUNCH_K((rank_tc_kernel<CAP_, BN_, ST_, 0, 4>));                                             \
        else TC_LAUNCH_K((rank_tc_kernel<CAP_, BN_, ST_, 0, 1>));                                                          \
    } while (0)
    const char* exp_s = getenv("YUE_RANK_EXP");
    const int exp_mode = exp_s ? atoi(exp_s) : 0;
    if (!atm && bn == 128 && cap == 64 && exp_mode == 1 && cl == 1) TC_LAUNCH_K((rank_tc_kernel<64, 128, 3, 1, 1>));
    else if (bn == 128 && cap == 64 && exp_mode == 1 && cl == 2) TC_LAUNCH_K((rank_tc_kernel<64, 128, 3, 1, 2>));
    else if (atm && exp_mode == 1 && cap == 64) TC_LAUNCH_K((rank_tc_kernel<64, 128, 3, 1, 1, true>));
    else if (atm && cap == 64 && stages == 3) TC_LAUNCH_K((rank_tc_kernel<64, 128, 3, 0, 1, true>));
    else if (atm && stages == 3) TC_LAUNCH_K((rank_tc_kernel<96, 128, 3, 0, 1, true>));
    else if (atm && cap == 64) TC_LAUNCH_K((rank_tc_kernel<64, 128, 2, 0, 1, true>));
    else if (atm) TC_LAUNCH_K((rank_tc_kernel<96, 128, 2, 0, 1, true>));
    else
    if (bn == 128) { if (cap == 64 && stages == 4) TC_LAUNCH(64, 128, 4); else if (cap == 64) TC_LAUNCH(64, 128, 3); else TC_LAUNCH(96, 128, 3); }
    else { if (cap == 64) TC_LAUNCH(64, 64, 2); else TC_LAUNCH(96, 64, 2); }
#undef TC_LAUNCH
#undef TC_LAUNCH_K
    ++launches;
    TC_CK(cudaGetLastError());

    int nfail = 0, nspill = 0;
    TC_CK(cudaMemcpyAsync(&nfail, st.fail_count, sizeof(int), cudaMemcpyDeviceToHost, stream));
    TC_CK(cudaMemcpyAsync(&nspill, st.ovf_next, sizeof(int), cudaMemcpyDeviceToHost, stream));
    TC_CK(cudaStreamSynchronize(stream));
    st.last_fail = nfail;
    st.last_spill = nspill;
    if (nfail > 0) {                      // rows with too many near-ties: exact kernel, then scatter back
        if ((size_t)nfail > st.fb_cap) {
            for (void* q : {(void*)st.fb_users, (void*)st.fb_ids, (void*)st.fb_scores}) if (q) cudaFree(q);
            st.fb_users = nullptr; st.fb_ids = nullptr; st.fb_scores = nullptr;
            TC_CK(cudaMalloc((void**)&st.fb_users, (size_t)nfail * sizeof(int32_t)));
            TC_CK(cudaMalloc((void**)&st.fb_ids, (size_t)nfail * 32 * sizeof(int32_t)));
            TC_CK(cudaMalloc((void**)&st.fb_scores, (size_t)nfail * 32 * sizeof(float)));
            st.fb_cap = (size_t)nfail;
        }
        tc_gather_users_kernel<<<(nfail + 255) / 256, 256, 0, stream>>>(d_users, st.fail_rows, nfail, st.fb_users);
        ++launches;
        if (int rc = fallback(st.fb_users, (int64_t)nfail, st.fb_ids, st.fb_scores)) return rc;
        tc_scatter_rows_kernel<<<(nfail * N + 255) / 256, 256, 0, stream>>>(st.fail_rows, nfail, N, st.fb_ids, st.fb_scores, d_ids, d_scores);
        ++launches;
        TC_CK(cudaGetLastError());
    }
    return 0;
#undef TC_CK
}

}  // namespace yue
