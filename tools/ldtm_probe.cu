// Microbenchmark: how fast can an SM read its tensor memory?  The claim to check (rank_tc_kernel): "TMEM reads
// run at 64 B per clock per SM, so draining a 128 x 128 fp32 accumulator tile takes 1024 cycles -- the floor of the fused
// ranking epilogue".  W warps (warp w reads the lanes of quarter w % 4; with 8 warps two warps share a quarter and split
// the columns) issue tcgen05.ld.32x32b back to back over all 512 columns, `iters` times; nothing else runs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ldtm_probe tools/ldtm_probe.cu && tools/ldtm_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD_REGS32(r) "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}"
#define OUT32(r, o) "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7]), \
                    "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]), "=r"(r[o + 12]), "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15]), \
                    "=r"(r[o + 16]), "=r"(r[o + 17]), "=r"(r[o + 18]), "=r"(r[o + 19]), "=r"(r[o + 20]), "=r"(r[o + 21]), "=r"(r[o + 22]), "=r"(r[o + 23]), \
                    "=r"(r[o + 24]), "=r"(r[o + 25]), "=r"(r[o + 26]), "=r"(r[o + 27]), "=r"(r[o + 28]), "=r"(r[o + 29]), "=r"(r[o + 30]), "=r"(r[o + 31])

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " LD_REGS32(r) ", [%32];" : OUT32(r, 0) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mode 0: issue all loads of a pass, one wait at the end of the pass (throughput of the read path)
// mode 1: wait after every load (latency-exposed, what a naive epilogue does)
// mode 2: like 0, plus a 31-step FMNMX chain per 32 columns on the PREVIOUS load's registers (the ranking epilogue's work)
__global__ void __launch_bounds__(256, 1) probe(int warps_active, int iters, int mode, unsigned long long* cycles, float* sink) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_slot;
    float acc = 0.f;
    unsigned long long t0 = 0, t1 = 0;
    if (warp < warps_active) {
        const int quarter = warp & 3, sharers = (warps_active + 3) / 4, part = warp >> 2;
        const uint32_t trow = base + ((uint32_t)(quarter * 32) << 16);
        // the columns this warp reads: with two warps per quarter each takes half of the 512
        const int c0 = 512 / sharers * part, c1 = 512 / sharers * (part + 1);
        uint32_t a[32], b[32];
        for (int x = 0; x < 32; ++x) a[x] = b[x] = 0;
        __syncwarp();
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            for (int c = c0; c < c1; c += 64) {
                ld32(trow + c, a);
                if (mode == 1) ld_wait();
                if (mode == 2) {
                    float m = __uint_as_float(b[0]);
#pragma unroll
                    for (int x = 1; x < 32; ++x) m = fmaxf(m, __uint_as_float(b[x]));
                    acc += m;
                }
                ld32(trow + c + 32, b);
                if (mode == 1) ld_wait();
                if (mode == 2) {
                    ld_wait();
                    float m = __uint_as_float(a[0]);
#pragma unroll
                    for (int x = 1; x < 32; ++x) m = fmaxf(m, __uint_as_float(a[x]));
                    acc += m;
                }
            }
            if (mode == 0) ld_wait();
        }
        ld_wait();
        t1 = clock64();
        for (int x = 0; x < 32; ++x) acc += __uint_as_float(a[x]) + __uint_as_float(b[x]);
    }
    if (threadIdx.x % 32 == 0 && warp < warps_active) cycles[blockIdx.x * 8 + warp] = t1 - t0;
    if (acc == 12345.f) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "r"(512u) : "memory");
}

int main() {
    unsigned long long* cyc;
    float* sink;
    cudaMalloc(&cyc, 148 * 8 * sizeof(unsigned long long));
    cudaMalloc(&sink, 4);
    const int iters = 2000;
    printf("tcgen05.ld.32x32b.x32 over the 512 TMEM columns, %d passes; bytes per clock PER SM = warps' bytes / the slowest warp's cycles\n", iters);
    const char* names[3] = {"back to back, one wait per pass", "wait after every load", "with a 31-step max per 32 columns"};
    for (int mode = 0; mode < 3; ++mode)
        for (int w : {1, 2, 4, 8}) {
            for (int grid : {1, 148}) {
                cudaMemset(cyc, 0, 148 * 8 * sizeof(unsigned long long));
                probe<<<grid, 256>>>(w, iters, mode, cyc, sink);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
                unsigned long long h[148 * 8];
                cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
                unsigned long long worst = 0;
                for (int i = 0; i < grid * 8; ++i) worst = h[i] > worst ? h[i] : worst;
                // every quarter's 32 lanes x 512 columns x 4 B are read once per pass by the warps that share the quarter
                const double bytes = (double)(w < 4 ? w : 4) * 32 * 512 * 4 * iters;
                printf("%-34s %d warp(s), %3d CTA(s): %9llu cycles  %6.1f B/clk/SM  (%5.1f cycles per 128-lane x 128-column fp32 tile%s)\n", names[mode], w, grid,
                       worst, bytes / worst, worst / (double)iters / 4.0 * (4.0 / (w < 4 ? w : 4)), w < 4 ? ", extrapolated to 4 quarters" : "");
            }
        }
    return 0;
}
