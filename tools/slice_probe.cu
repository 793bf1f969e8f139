// Which 256-byte granules share an L2 slice with granule 0?  All warps RED a full 256-byte row into
// granule 0 (even warps) or granule g (odd warps); a slice serves ~1 sector per clock, so the pair
// takes twice as long when both granules live in the same slice.  Prints the conflicting g.
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
__global__ void k(float* buf, size_t gB, int iters) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float* p = buf + ((warp & 1) ? gB * 64 : 0) + lane * 2;
    for (int it = 0; it < iters; ++it)
        asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %1};" :: "l"(p), "f"(1.0f) : "memory");
}
int main(int argc, char** argv) {
    const int G = argc > 1 ? atoi(argv[1]) : 2048;
    float* buf; cudaMalloc(&buf, (size_t)64 << 20); cudaMemset(buf, 0, (size_t)64 << 20);
    printf("base address %p (offset in 64 MB: %zu)\n", (void*)buf, (size_t)buf & ((64u << 20) - 1));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<float> ms(G);
    for (int g = 0; g < G; ++g) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0); k<<<148, 320>>>(buf, (size_t)g, 100); cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        cudaEventElapsedTime(&ms[g], e0, e1);
    }
    printf("same granule (g=0): %.3f ms; g=1: %.3f ms; g=2: %.3f ms\n", ms[0], ms[1], ms[2]);
    const float thr = 0.75f * ms[0];
    printf("granules conflicting with granule 0 (time > %.3f ms):", thr);
    int n = 0;
    for (int g = 1; g < G; ++g) if (ms[g] > thr) { printf(" %d", g); ++n; }
    printf("\n%d of %d\n", n, G - 1);
    // xor-linearity check on a few pairs
    return 0;
}
