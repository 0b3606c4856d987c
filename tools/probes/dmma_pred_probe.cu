// Probe: does a predicated-off DMMA (mma.sync m8n8k4 f64) still occupy the FP64 tensor pipe?
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma_pred(double& c0, double& c1, double a, double b, int on)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %4, 0;\n\t@p mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n\t}"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b), "r"(on));
}
// mode 0: 32 plain DMMAs per iteration; mode 1: 16 plain + 16 predicated (on = runtime 0); mode 2: 16 plain only;
// mode 3: 16 plain + 16 behind a warp-uniform branch not taken
__global__ void k(double* out, int iters, int mode, int on)
{
    double acc[32][2];
    for (int i = 0; i < 32; ++i) acc[i][0] = acc[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = 1.0 + blockIdx.x * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma(acc[i][0], acc[i][1], a, b);
        if (mode == 0) {
#pragma unroll
            for (int i = 16; i < 32; ++i) dmma(acc[i][0], acc[i][1], a, b);
        } else if (mode == 1) {
#pragma unroll
            for (int i = 16; i < 32; ++i) dmma_pred(acc[i][0], acc[i][1], a, b, on);
        } else if (mode == 3) {
            if (on) {
#pragma unroll
                for (int i = 16; i < 32; ++i) dmma(acc[i][0], acc[i][1], a, b);
            }
        }
    }
    double s = 0; for (int i = 0; i < 32; ++i) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    double* out; cudaMalloc(&out, 148 * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 4; ++mode)
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            k<<<148, 256>>>(out, iters, mode, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double nd = (mode == 0 ? 32.0 : 16.0) * iters * 8 * 148;
            printf("mode %d rep %d: %.3f ms, executed DMMA rate %.2f TFLOP/s\n", mode, rep, ms, nd * 512 / ms / 1e9);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
