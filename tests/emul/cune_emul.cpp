// Test infrastructure: the device code of yue_b200/csrc/cune_sgd.cuh (K8) compiled for the HOST with a one-lane "warp"
// (W = 1: a lane owns every column, warp shuffles are identities), so that the statement order, the aliasing of j and k,
// the fused Philox draws and the loss of the kernel can be checked against tests/golden/cune_small.npz on a box without a
// GPU.  It checks the kernel's TEXT; the 32-lane reductions, the memory system and the launch are what the -m gpu tests
// check.  Built by tests/test_zz_cune.py with g++ -ffp-contract=off; never loaded by the product.
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
struct EmulIdx { unsigned x = 0, y = 0, z = 0; };
static EmulIdx threadIdx;
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
static inline void __syncwarp() {}
#define __expf(x) expf(x)
static inline float __fdividef(float a, float b) { return a / b; }
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }

#include "../../yue_b200/csrc/cune_sgd.cuh"

extern "C" int cune_emul_epoch(float* P, float* Q, int ld, int k, int64_t m, int64_t n, const int64_t* ev_indptr,
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
    if (serial) yue::cune_sgd_kernel<16, yue::kSerial, 1>(cp);
    else yue::cune_sgd_kernel<16, yue::kAtomic, 1>(cp);
    *loss_out = loss;
    *users_out = ctr[1];
    return 0;
}
