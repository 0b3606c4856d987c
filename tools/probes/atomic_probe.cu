// Probe: throughput of FP64 RED.ADD for a 128x128 tile scattered through a monotone station map,
// (A) in the DMMA fragment layout (each lane its own 32-B sector) vs (B) row-wise, lanes along columns.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__global__ void frag_layout(double* C, const int* rowmap, int ldc, int ntile_side)
{
    const int tile = blockIdx.x; const int tm = tile / ntile_side, tn = tile % ntile_side;
    if (tn > tm) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;
    for (int i = 0; i < 8; ++i) {
        const int r = tm * 128 + wm + 16 * (i >> 1) + 2 * g + (i & 1);
        const long dr = 3l * rowmap[r / 3] + r % 3;
        for (int j = 0; j < 4; ++j)
            for (int e = 0; e < 2; ++e) {
                const int c = tn * 128 + wn + 16 * (j >> 1) + 2 * (2 * t + e) + (j & 1);
                if (r < c) continue;
                const long dc = 3l * rowmap[c / 3] + c % 3;
                atomicAdd(C + dr * ldc + dc, 1.0);
            }
    }
}
__global__ void row_layout(double* C, const int* rowmap, int ldc, int ntile_side)
{
    const int tile = blockIdx.x; const int tm = tile / ntile_side, tn = tile % ntile_side;
    if (tn > tm) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int rr = warp; rr < 128; rr += 8) {
        const int r = tm * 128 + rr;
        const long dr = 3l * rowmap[r / 3] + r % 3;
        for (int cc = lane; cc < 128; cc += 32) {
            const int c = tn * 128 + cc;
            if (r < c) continue;
            const long dc = 3l * rowmap[c / 3] + c % 3;
            atomicAdd(C + dr * ldc + dc, 1.0);
        }
    }
}
int main()
{
    const int side = 24, n = side * 128, nst = n / 3;          // 3072 x 3072 source, scattered into a wider target
    std::vector<int> map(nst);
    srand(1); int pos = 0;
    for (int i = 0; i < nst; ++i) { map[i] = pos; pos += 1 + (rand() % 4 == 0 ? rand() % 5 : 0); }
    const int ldc = 3 * pos + 2; const size_t bytes = (size_t)ldc * 3 * pos * 8;
    double* C; int* d_map; cudaMalloc(&C, bytes); cudaMemset(C, 0, bytes); cudaMalloc(&d_map, nst * 4);
    cudaMemcpy(d_map, map.data(), nst * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double natom = (double)n * (n + 1) / 2;
    for (int mode = 0; mode < 2; ++mode)
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) frag_layout<<<side * side, 256>>>(C, d_map, ldc, side);
            else row_layout<<<side * side, 256>>>(C, d_map, ldc, side);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("%s rep %d: %.3f ms, %.1f G atomics/s (target %.0f MB)\n", mode ? "row " : "frag", rep, ms, natom / ms / 1e6, bytes / 1e6);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
