// Microbenchmark: throughput of 256-byte row REDs (vector float atomics) into R rows, to find what
// bounds the hot rows of the SGD kernel (per address? per L2 slice? per request?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/red_probe tools/red_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int V>
__device__ __forceinline__ void red(float* p, float v) {
    if (V == 2) asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %1};" :: "l"(p), "f"(v) : "memory");
    else if (V == 4) asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %1, %1, %1};" :: "l"(p), "f"(v) : "memory");
    else asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}

// every warp: iters REDs of one 64-float row; row = (warp * 7 + it * step) % rows; lanes_per_row * V = 64
template <int V>
__global__ void k(float* buf, int rows, size_t stride, int iters, int step, int also_load) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    constexpr int LPR = 64 / V;                 // lanes per row
    if (lane >= LPR) return;
    float acc = 0.f;
    unsigned r = warp * 7u;
    for (int it = 0; it < iters; ++it) {
        r += step;
        float* p = buf + (size_t)(r % rows) * stride + lane * V;
        if (also_load) acc += __ldcg(p);
        red<V>(p, 1.0f + acc * 0.f);
    }
    if (acc == 123.f) buf[0] = acc;
}

int main() {
    float* buf;
    const size_t bytes = (size_t)1 << 30;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148, block = 320, warps = grid * block / 32;
    auto run = [&](const char* name, int V, int rows, size_t stride, int iters, int step, int load) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (V == 2) k<2><<<grid, block>>>(buf, rows, stride, iters, step, load);
            else if (V == 4) k<4><<<grid, block>>>(buf, rows, stride, iters, step, load);
            else k<1><<<grid, block>>>(buf, rows, stride, iters, step, load);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double n = (double)warps * iters;
        printf("%-28s V=%d rows=%7d stride=%6zu load=%d : %8.3f ms  %7.2f ns/row-RED (serial)  %8.1f M row-RED/s  per-row chain %.2f ns\n",
               name, V, rows, stride, load, ms, ms * 1e6 / n, n / ms / 1e3, ms * 1e6 / (n / rows));
    };
    const int it = 2000;
    for (int V : {2, 4, 1}) {
        run("same row", V, 1, 64, it, 1, 0);
        run("2 rows contiguous", V, 2, 64, it, 1, 0);
        run("4 rows contiguous", V, 4, 64, it, 1, 0);
        run("8 rows contiguous", V, 8, 64, it, 1, 0);
        run("16 rows contiguous", V, 16, 64, it, 1, 0);
        run("64 rows contiguous", V, 64, 64, it, 1, 0);
        run("1024 rows contiguous", V, 1024, 64, it, 1, 0);
        run("200K rows contiguous", V, 200000, 64, it, 1, 0);
        run("200K rows, step 7919", V, 200000, 64, it, 7919, 0);
        run("4 rows stride 4 KB+256", V, 4, 1024 + 64, it, 1, 0);
        run("8 rows stride 4 KB+256", V, 8, 1024 + 64, it, 1, 0);
        run("4 rows stride 1 MB+256", V, 4, 262144 + 64, it, 1, 0);
        run("8 rows stride 1 MB+256", V, 8, 262144 + 64, it, 1, 0);
        run("8 rows stride 768 B", V, 8, 192, it, 1, 0);
        run("200K rows, step 7919 + load", V, 200000, 64, it, 7919, 1);
    }
    return 0;
}
