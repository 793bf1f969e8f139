// Test infrastructure: the device code of yue_b200/csrc/cune_sgd.cuh (K8) compiled for the HOST with a one-lane "warp"
// (W = 1: a lane owns every column, warp shuffles are identities), so that the statement order, the aliasing of j and k,
// the fused Philox draws and the loss of the kernel can be checked against tests/golden/cune_small.npz on a box without a
// GPU.  With -DEMUL_LANES=8 the warp is eight host threads in lockstep instead (a shuffle is an exchange through a table
// between two barriers): the butterflies, the lane-per-event draws and their broadcast, the column ownership and the
// cross-lane visibility rules run as written, for a warp of 8.  It checks the kernel's TEXT; the memory system, the
// 32-lane width and the launch are what the -m gpu tests check.  Built by tests/test_zz_cune.py with g++
// -ffp-contract=off -fvisibility=hidden -Wl,-Bsymbolic; never loaded by the product.
#define YUE_CUNE_HOST_EMUL 1
#include <cmath>
#include <cstdint>
#include <cstring>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#ifndef EMUL_LANES
#define EMUL_LANES 1
#endif
struct EmulIdx { unsigned x = 0, y = 0, z = 0; };
#if EMUL_LANES == 1
static EmulIdx threadIdx;
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
static inline void __syncwarp() {}
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
#else
#include <barrier>
#include <mutex>
#include <thread>
static thread_local EmulIdx threadIdx;
static std::barrier<>* g_bar = nullptr;
static uint64_t g_xchg[EMUL_LANES];
static std::mutex g_atomic;
template <class T> static inline T emul_exchange(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle operand");
    uint64_t raw = 0, got;
    memcpy(&raw, &v, sizeof(T));
    g_xchg[threadIdx.x] = raw;
    g_bar->arrive_and_wait();
    got = g_xchg[src & (EMUL_LANES - 1)];
    g_bar->arrive_and_wait();
    T out;
    memcpy(&out, &got, sizeof(T));
    return out;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emul_exchange(v, src); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emul_exchange(v, (int)threadIdx.x ^ m); }
static inline void __syncwarp() { g_bar->arrive_and_wait(); }
template <class T> static inline T atomicAdd(T* p, T v) { std::lock_guard<std::mutex> l(g_atomic); T o = *p; *p = o + v; return o; }
#endif
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
#define __expf(x) expf(x)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }

struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
namespace yue {                      // the three row accessors of bpr_sgd.cuh (PTX there)
static inline float4 ld_row(const float* p) { return float4{p[0], p[1], p[2], p[3]}; }
static inline void st_row(float* p, float4 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w; }
static inline void red_row(float* p, float4 v) { atomicAdd(p, v.x); atomicAdd(p + 1, v.y); atomicAdd(p + 2, v.z); atomicAdd(p + 3, v.w); }
}  // namespace yue

#include "../../yue_b200/csrc/cune_sgd.cuh"

// Built with -fvisibility=hidden -Wl,-Bsymbolic: libyue_b200.so, when it is loaded in the same process, exports host stubs
// with the very names of the instantiations used here (yue::cune_sgd_kernel<1, MODE, 32>), and they must not be bound to.
extern "C" __attribute__((visibility("default"))) int cune_emul_epoch(float* P, float* Q, int ld, int k, int64_t m, int64_t n, const int64_t* ev_indptr,
                               const int32_t* ev_items, const int64_t* uq_indptr, const int32_t* uq_items,
                               const int64_t* ip_indptr, const int32_t* ip_items, uint64_t seed, uint32_t epoch, double lr,
                               double regU, double regI, double s, int serial, double* loss_out, uint64_t* users_out,
                               const int32_t* hot_items, int64_t event_base, const int64_t* ev_delta, int64_t chunk) {
    if (ld > 16) return 1;
    unsigned long long ctr[2] = {0, 0};
    double loss = 0.0;
    yue::CuneParams cp{};
    cp.P = P; cp.Q = Q; cp.ld = ld; cp.k = k; cp.m = m; cp.n = n;
    cp.ev_indptr = ev_indptr; cp.ev_items = ev_items; cp.hot_items = hot_items;
    cp.uq_indptr = uq_indptr; cp.uq_items = uq_items; cp.ip_indptr = ip_indptr; cp.ip_items = ip_items;
    cp.seed = seed; cp.epoch = epoch; cp.event_base = event_base; cp.ev_delta = ev_delta;
    cp.lr = lr; cp.inv_s = 1.0 / s; cp.regU = regU; cp.regI = regI;
    cp.c_u = (float)(lr * regU); cp.c_i = (float)(lr * regI);
    std::vector<int64_t> items;
    yue::cune_plan_items(m, ev_indptr, serial ? 0 : chunk, items);
    cp.items = items.data(); cp.n_work = (int64_t)items.size() / yue::kCuneItemWords;
    cp.cursor = &ctr[0]; cp.users_done = &ctr[1]; cp.loss = &loss;
    constexpr int kNC = (4 + EMUL_LANES - 1) / EMUL_LANES;      // 16-byte chunks per lane for ld <= 16
#if EMUL_LANES == 1
    if (serial) yue::cune_sgd_kernel<kNC, yue::kSerial, 1>(cp);
    else yue::cune_sgd_kernel<kNC, yue::kAtomic, 1>(cp);
#else
    std::barrier<> bar(EMUL_LANES);
    g_bar = &bar;
    std::thread lanes[EMUL_LANES];
    for (int l = 0; l < EMUL_LANES; ++l)
        lanes[l] = std::thread([&cp, serial, l]() {
            threadIdx.x = (unsigned)l;
            if (serial) yue::cune_sgd_kernel<kNC, yue::kSerial, EMUL_LANES>(cp);
            else yue::cune_sgd_kernel<kNC, yue::kAtomic, EMUL_LANES>(cp);
        });
    for (auto& t : lanes) t.join();
    g_bar = nullptr;
#endif
    *loss_out = loss;
    *users_out = ctr[1];
    return 0;
}
