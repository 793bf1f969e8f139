// Does an L1-cached load (ld.ca) of a line that the same SM keeps RED-ing see the updates, and what does
// it cost?  All warps: loop { load a 256-byte row with the given cache operator; RED +1 into it }.
// Reports ns per iteration and how far behind the last loaded value is from the final value in memory.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>   // 0 = ld.cg (L2), 1 = ld.ca (L1), 2 = ld.cv (volatile), 3 = ca with a cv refresh every 8th
__global__ void k(float* buf, size_t stride_f, int nsec, int iters, float* last) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if ((lane >> 2) >= nsec) return;
    float* p = buf + (size_t)(lane >> 2) * stride_f + (lane & 3) * 2;
    float v = 0.f;
    for (int it = 0; it < iters; ++it) {
        float2 t;
        const bool refresh = MODE == 2 || (MODE == 3 && (it & 7) == 0);
        if (MODE == 0) asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(t.x), "=f"(t.y) : "l"(p));
        else if (refresh) asm volatile("ld.global.cv.v2.f32 {%0, %1}, [%2];" : "=f"(t.x), "=f"(t.y) : "l"(p));
        else asm volatile("ld.global.ca.v2.f32 {%0, %1}, [%2];" : "=f"(t.x), "=f"(t.y) : "l"(p));
        v = t.x;
        asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %1};" :: "l"(p), "f"(1.0f + v * 0.f) : "memory");
    }
    if (lane == 0) last[warp] = v;
}
int main() {
    float* buf; cudaMalloc(&buf, 64 << 20);
    float* last; cudaMalloc(&last, 1480 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000, warps = 1480;
    for (int layout = 0; layout < 2; ++layout)
        for (int mode = 0; mode < 4; ++mode) {
            cudaMemset(buf, 0, 64 << 20);
            const size_t stride_f = layout ? 65 * 64 : 8;       // spread (8 slices) or contiguous row
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148, 320>>>(buf, stride_f, 8, iters, last);
            if (mode == 1) k<1><<<148, 320>>>(buf, stride_f, 8, iters, last);
            if (mode == 2) k<2><<<148, 320>>>(buf, stride_f, 8, iters, last);
            if (mode == 3) k<3><<<148, 320>>>(buf, stride_f, 8, iters, last);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            float h[1480], fin; cudaMemcpy(h, last, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(&fin, buf, 4, cudaMemcpyDeviceToHost);
            double mn = 1e30, mean = 0; for (int w = 0; w < warps; ++w) { mean += h[w]; if (h[w] < mn) mn = h[w]; }
            printf("%s row, %s: %7.3f ms  %5.2f ns per load+RED; final %.0f, last value loaded: mean %.0f min %.0f (of %d total REDs)\n",
                   layout ? "spread    " : "contiguous", mode == 0 ? "ld.cg          " : mode == 1 ? "ld.ca          " : mode == 2 ? "ld.cv          " : "ld.ca, cv each 8",
                   ms, ms * 1e6 / ((double)warps * iters), fin, mean / warps, mn, warps * iters);
        }
    return 0;
}
