// Microbenchmark 2: one logical 256-byte row whose eight 32-byte sectors sit `sstride` bytes apart;
// all warps RED (v2.f32, 4 lanes per sector) into it.  sectors=1 -> a single 32-byte address.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* buf, int nsec, size_t sstride_f, int rows, size_t rstride_f, int iters, int load) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int sec = lane >> 2;
    if (sec >= nsec) return;
    unsigned r = warp * 7u;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        ++r;
        float* p = buf + (size_t)(r % rows) * rstride_f + sec * sstride_f + (lane & 3) * 2;
        if (load) { float2 t = __ldcg((const float2*)p); acc += t.x; }
        asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %1};" :: "l"(p), "f"(1.0f + acc * 0.f) : "memory");
    }
    if (acc == 123.f) buf[0] = acc;
}
int main() {
    float* buf; const size_t bytes = (size_t)1 << 30;
    cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148, block = 320, warps = grid * block / 32, iters = 2000;
    auto run = [&](int nsec, size_t sstride, int rows, size_t rstride, int load) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            k<<<grid, block>>>(buf, nsec, sstride / 4, rows, rstride / 4, iters, load);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double n = (double)warps * iters;
        printf("sectors=%d sector-stride=%7zu B rows=%d row-stride=%7zu B load=%d : %7.3f ms  %6.2f ns per row-RED  (%.2f ns per row-RED per row)\n",
               nsec, sstride, rows, rstride, load, ms, ms * 1e6 / n, ms * 1e6 / n * rows);
    };
    for (int load : {0, 1}) {
        run(8, 32, 1, 256, load);          // contiguous row (baseline 4.08 ns)
        run(1, 32, 1, 256, load);          // a single sector
        run(2, 32, 1, 256, load);
        run(4, 32, 1, 256, load);
        run(8, 256, 1, 4096, load);        // sectors 256 B apart
        run(8, 288, 1, 4096, load);
        run(8, 768, 1, 8192, load);
        run(8, 1024 + 256, 1, 16384, load);
        run(8, 4096 + 256, 1, 65536, load);
        run(8, 65536 + 256, 1, 1 << 20, load);
        run(8, 768, 4, 8192 + 32, load);   // 4 such rows
        run(8, 32, 4, 768, load);          // 4 contiguous-sector rows in 4 slices
    }
    return 0;
}
