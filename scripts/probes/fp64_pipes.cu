// Do DMMA (mma.sync.m8n8k4.f64) and DFMA share one datapath on B200?  Three kernels, same grid (148 x 4 CTAs x 8 warps):
// all warps DFMA chains, all warps DMMA chains, half/half.  Reports instruction rates; if the mixed kernel sustains
// both rates at once the pipes are independent.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, int iters)
{
    const int warp = threadIdx.x >> 5;
    const bool do_mma = MODE == 1 || (MODE == 2 && (warp & 1));
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - 1e-9;
    double d[8][2];
    for (int i = 0; i < 8; ++i) { d[i][0] = i; d[i][1] = -i; }
    double f[16];
    for (int i = 0; i < 16; ++i) f[i] = i * 0.5;
    if (do_mma) {
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[i][0]), "+d"(d[i][1]) : "d"(a), "d"(b));
    } else {
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fma(f[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1];
    for (int i = 0; i < 16; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> float run(double* out, int iters)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 256>>>(out, 16);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 256>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main()
{
    double* out; cudaMalloc(&out, 148 * 4 * 256 * 8);
    const int iters = 20000;
    const double warps = 148.0 * 4 * 8;
    float t0 = run<0>(out, iters), t1 = run<1>(out, iters), t2 = run<2>(out, iters);
    // per-SM per-clock rates at 1.965 GHz
    const double clk = 1.965e9;
    printf("DFMA only : %.3f ms  -> %.2f warp-DFMA / clk / SM\n", t0, warps * iters * 16 / (t0 * 1e-3) / clk / 148);
    printf("DMMA only : %.3f ms  -> %.2f DMMA / clk / SM  (%.1f TFLOP/s)\n", t1, warps * iters * 8 / (t1 * 1e-3) / clk / 148, warps * iters * 8 * 512 / (t1 * 1e-3) / 1e12);
    printf("half/half : %.3f ms  -> %.2f warp-DFMA + %.2f DMMA / clk / SM\n", t2, warps / 2 * iters * 16 / (t2 * 1e-3) / clk / 148, warps / 2 * iters * 8 / (t2 * 1e-3) / clk / 148);
    printf("if the pipes were one, half/half would take (t0 + t1) / 2 = %.3f ms; independent pipes: max(t0, t1) / 2 = %.3f ms\n", (t0 + t1) / 2, (t0 > t1 ? t0 : t1) / 2);
    return 0;
}
