// Microbenchmark 3: candidate layouts of the hot-row table.  Every warp iteration reads and REDs one
// 256-byte logical row chosen with zipf(1) frequency among NS hot slots (like the hot positives of C2).
//   layout 0: rows contiguous (256 B each, like Q)          layout 1: sector c of slot s at c*PLANE + s*256
//   layout 2: as 1 with slot s in granule (s/2)*4 + s%2 (address bit 9 stays 0: the slice hash ignores it)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
__global__ void k(float* buf, const int* tab, int layout, size_t plane_f, int iters) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t lane_spread = (size_t)(lane >> 2) * plane_f + (lane & 3) * 2, lane_cont = lane * 2;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        const int s = tab[(warp * 131 + it * 7) & 4095];
        float* p = layout == 2 ? buf + (size_t)(((s >> 1) << 2) | (s & 1)) * 64 + lane_spread : layout ? buf + (size_t)s * 64 + lane_spread : buf + (size_t)s * 64 + lane_cont;
        float2 t = __ldcg((const float2*)p); acc += t.x;
        asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %1};" :: "l"(p), "f"(1.0f + acc * 0.f) : "memory");
    }
    if (acc == 123.f) buf[0] = acc;
}
int main() {
    float* buf; cudaMalloc(&buf, 64 << 20); cudaMemset(buf, 0, 64 << 20);
    int* dtab; cudaMalloc(&dtab, 4096 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148, block = 320, warps = grid * block / 32, iters = 2000;
    for (int NS : {1, 4, 9, 16, 32}) {
        std::vector<int> tab(4096);
        double H = 0; for (int s = 0; s < NS; ++s) H += 1.0 / (s + 1);
        int pos = 0;
        for (int s = 0; s < NS; ++s) { int cnt = (int)std::round(4096.0 / (s + 1) / H); for (int c = 0; c < cnt && pos < 4096; ++c) tab[pos++] = s; }
        while (pos < 4096) tab[pos++] = 0;
        for (int i = 4095; i > 0; --i) std::swap(tab[i], tab[rand() % (i + 1)]);
        cudaMemcpy(dtab, tab.data(), 4096 * 4, cudaMemcpyHostToDevice);
        auto run = [&](int layout, size_t plane_bytes) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0); k<<<grid, block>>>(buf, dtab, layout, plane_bytes / 4, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("slots=%2d layout=%d plane=%6zu B : %7.3f ms  %5.2f ns per touch (load+RED of one row)\n", NS, layout, plane_bytes, ms, ms * 1e6 / ((double)warps * iters));
        };
        run(0, 0);
        for (size_t plane : {(size_t)65 * 256, (size_t)67 * 256, (size_t)64 * 256, (size_t)129 * 256, (size_t)65 * 256 + 32})
            run(1, plane);
        for (size_t plane : {(size_t)65 * 256, (size_t)67 * 256, (size_t)129 * 256})
            run(2, plane);
    }
    return 0;
}
