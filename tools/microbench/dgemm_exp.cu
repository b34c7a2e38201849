// Stand-alone timing harness for k_dgemm (Gram shape): nvcc [-DGM_EXP_NOBAR] [-DGM_EXP_NOLOAD] ... (results are wrong with the
// experiment macros; they bound what the barrier / the load issue cost).
#include <cstdio>
#include <vector>
#include "../../reducedbasismethods.jl_b200/csrc/dgemm_kernels.cuh"
using namespace rbm;
#ifndef FULL
#define FULL 1
#endif
__global__ void fill(double *p, size_t n) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { unsigned long long z = (i + 1) * 0x9E3779B97F4A7C15ull; z ^= z >> 29; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 32; p[i] = FULL ? (double)(z >> 11) / 9007199254740992.0 - 0.5 : (double)(z & 1023) / 1024.0 - 0.5; } }
int main(int argc, char **argv)
{
    const int64_t K = argc > 1 ? atoll(argv[1]) : 400000, N = argc > 2 ? atoll(argv[2]) : 2000;
    const int nsplit = argc > 3 ? atoi(argv[3]) : 13;
    double *X; cudaMalloc(&X, sizeof(double) * K * N); fill<<<1024, 256>>>(X, (size_t)K * N);
    std::vector<const double *> h(N); for (int64_t c = 0; c < N; ++c) h[c] = X + c * K;
    const double **d; cudaMalloc(&d, sizeof(double *) * N); cudaMemcpy(d, h.data(), sizeof(double *) * N, cudaMemcpyHostToDevice);
    const int tm = (int)((N + GM_BM - 1) / GM_BM), ntiles = tm * (tm + 1) / 2;
    int64_t kps = (K + nsplit - 1) / nsplit; kps = (kps + GM_BK - 1) / GM_BK * GM_BK;
    double *W; cudaMalloc(&W, sizeof(double) * nsplit * N * N);
    GemmArgs a{}; a.acol = d; a.bcol = d; a.M = N; a.N = N; a.K = K; a.k_per_split = kps; a.out = W; a.ldo = N; a.split_stride = N * N; a.tiles_m = tm; a.same_ab = 1;
    const size_t smem = sizeof(double) * GM_STAGES * (GM_A_ELEMS_T + GM_B_ELEMS) + sizeof(double *) * (GM_BM + GM_BN);
    cudaFuncSetAttribute(k_dgemm<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_dgemm_mb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GM_MB_SMEM);
    const int bulk = argc > 4 ? atoi(argv[4]) : 0;
    cudaFuncSetAttribute(k_dgemm_pmb<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gm_pmb_smem<true>());
    double *W2; cudaMalloc(&W2, sizeof(double) * nsplit * N * N);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dim3 grid(ntiles, (unsigned)((K + kps - 1) / kps));
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (bulk == 2) { const int nitems = (int)(grid.x * grid.y); k_dgemm_pmb<true><<<nitems < 148 ? nitems : 148, GM_THREADS, gm_pmb_smem<true>()>>>(a, (int)grid.x, nitems); }
        else if (bulk) k_dgemm_mb<<<grid, GM_THREADS, GM_MB_SMEM>>>(a); else k_dgemm<true, true><<<grid, GM_THREADS, smem>>>(a);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("K=%lld N=%lld grid %u x %u: %.2f ms  useful %.2f TF  executed %.2f TF  (%s)\n", (long long)K, (long long)N, grid.x, grid.y, ms,
               (double)K * N * (N + 1) / ms / 1e9, 2.0 * K * ntiles * GM_BM * GM_BN / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    if (bulk) {   // compare against the cp.async kernel (lower-triangular tiles only)
        GemmArgs a2 = a; a2.out = W2;
        k_dgemm<true, true><<<grid, GM_THREADS, smem>>>(a2);
        std::vector<double> h1((size_t)nsplit * N * N), h2(h1.size());
        cudaMemcpy(h1.data(), W, sizeof(double) * h1.size(), cudaMemcpyDeviceToHost);
        cudaMemcpy(h2.data(), W2, sizeof(double) * h2.size(), cudaMemcpyDeviceToHost);
        double md = 0, mx = 0; size_t cnt = 0;
        for (unsigned sp = 0; sp < grid.y; ++sp) for (int64_t j = 0; j < N; ++j) for (int64_t i = j / GM_BN * GM_BN; i < N; ++i) {
            size_t q = (size_t)sp * N * N + j * N + i; double d1 = h1[q] - h2[q]; if (d1 < 0) d1 = -d1; if (d1 > md) md = d1;
            double v = h2[q] < 0 ? -h2[q] : h2[q]; if (v > mx) mx = v; ++cnt; }
        printf("bulk vs cp.async: max |diff| %.3e (max |value| %.3e over %zu entries) %s\n", md, mx, cnt, cudaGetErrorString(cudaDeviceSynchronize()));
    }
    return 0;
}
