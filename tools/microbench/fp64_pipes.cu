// Are the FP64 FMA pipe and the FP64 tensor (DMMA) sub-pipe separate issue resources on B200?
// Three kernels with register-only operands: DFMA only, DMMA only, both interleaved.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu && ./fp64_pipes
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE> __global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b)
{
    double c[16], f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i] = threadIdx.x * 1e-3 + i; f[i] = i * 0.5; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE & 1) dmma(c[2 * i], c[2 * i + 1], a, b);
            if (MODE & 2) { f[2 * i] = fma(f[2 * i], a, b); f[2 * i + 1] = fma(f[2 * i + 1], a, b); }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, double *d)
{
    const int iters = 20000, blocks = 148 * 2;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warps = blocks * 8.0;
    double dmma_fl = (MODE & 1) ? warps * iters * 8.0 * 512.0 : 0.0;        // 8 DMMA x 256 FMA x 2
    double dfma_fl = (MODE & 2) ? warps * 32.0 * iters * 16.0 * 2.0 : 0.0;  // 16 DFMA per thread
    printf("%-12s %.3f ms  DMMA %.2f TFLOP/s  DFMA %.2f TFLOP/s  total %.2f\n", name, ms, dmma_fl / ms / 1e9, dfma_fl / ms / 1e9,
           (dmma_fl + dfma_fl) / ms / 1e9);
}
int main()
{
    double *d; cudaMalloc(&d, 148 * 2 * 256 * sizeof(double));
    run<1>("DMMA only", d); run<2>("DFMA only", d); run<3>("DMMA + DFMA", d);
    return 0;
}
