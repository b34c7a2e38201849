// DMMA.8x8x4 throughput as a function of warps per SM and independent accumulator chains per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_occupancy dmma_occupancy.cu && ./dmma_occupancy
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ILP> __global__ void k(double *out, int iters, double a, double b)
{
    double c[2 * ILP];
#pragma unroll
    for (int i = 0; i < 2 * ILP; ++i) c[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 32 / ILP; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) dmma(c[2 * i], c[2 * i + 1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 2 * ILP; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> void run(double *d, int warps)
{
    const int iters = 5000, blocks = 148;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP><<<blocks, warps * 32>>>(d, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0); k<ILP><<<blocks, warps * 32>>>(d, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = (double)blocks * warps * iters * 32.0 * 512.0;
    printf("warps/SM %2d  chains %2d  %.2f TFLOP/s\n", warps, ILP, fl / ms / 1e9);
}
int main()
{
    double *d; cudaMalloc(&d, 148 * 1024 * sizeof(double));
    for (int w : {4, 8, 16, 32}) { run<1>(d, w); run<2>(d, w); run<4>(d, w); run<8>(d, w); run<32>(d, w); }
    return 0;
}
