// Probe: is mma.sync.m8n8k4.f64 bit-identical to a sequential chain of IEEE FMAs in increasing k
// (d = fma(a3,b3, fma(a2,b2, fma(a1,b1, fma(a0,b0,c)))))?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

__global__ void k(const double* A, const double* B, const double* C, double* D, int n_tiles)
{
    const int lane = threadIdx.x & 31;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const double* a = A + (size_t)t * 32;   // 8x4 row-major
        const double* b = B + (size_t)t * 32;   // 4x8 "col" operand: element (k, n) at b[n*4 + k]
        const double* c = C + (size_t)t * 64;   // 8x8 row-major
        double av = a[(lane >> 2) * 4 + (lane & 3)];
        double bv = b[(lane >> 2) * 4 + (lane & 3)];
        double c0 = c[(lane >> 2) * 8 + (lane & 3) * 2], c1 = c[(lane >> 2) * 8 + (lane & 3) * 2 + 1];
        double d0, d1;
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                     : "=d"(d0), "=d"(d1) : "d"(av), "d"(bv), "d"(c0), "d"(c1));
        D[(size_t)t * 64 + (lane >> 2) * 8 + (lane & 3) * 2] = d0;
        D[(size_t)t * 64 + (lane >> 2) * 8 + (lane & 3) * 2 + 1] = d1;
    }
}

int main()
{
    const int T = 20000;
    double *A, *B, *C, *D;
    cudaMallocManaged(&A, T * 32 * 8); cudaMallocManaged(&B, T * 32 * 8);
    cudaMallocManaged(&C, T * 64 * 8); cudaMallocManaged(&D, T * 64 * 8);
    srand(1);
    auto rnd = []() { return (rand() / (double)RAND_MAX - 0.5) * exp((rand() % 40 - 20) * 0.5); };
    for (int i = 0; i < T * 32; ++i) { A[i] = rnd(); B[i] = rnd(); }
    for (int i = 0; i < T * 64; ++i) C[i] = rnd();
    k<<<148, 32>>>(A, B, C, D, T);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 1; }
    long fwd = 0, rev = 0, pair = 0, tot = 0;
    for (int t = 0; t < T; ++t)
        for (int m = 0; m < 8; ++m)
            for (int n = 0; n < 8; ++n) {
                const double* a = A + (size_t)t * 32 + m * 4;
                const double* b = B + (size_t)t * 32 + n * 4;
                double c = C[(size_t)t * 64 + m * 8 + n], d = D[(size_t)t * 64 + m * 8 + n];
                double f = c; for (int kk = 0; kk < 4; ++kk) f = fma(a[kk], b[kk], f);
                double r = c; for (int kk = 3; kk >= 0; --kk) r = fma(a[kk], b[kk], r);
                double p = fma(a[1], b[1], a[0] * b[0]) + fma(a[3], b[3], a[2] * b[2]) + c;
                fwd += (f == d); rev += (r == d); pair += (p == d); ++tot;
            }
    printf("elements %ld: equal to forward FMA chain %ld, reverse chain %ld, pairwise %ld\n", tot, fwd, rev, pair);
    return 0;
}
