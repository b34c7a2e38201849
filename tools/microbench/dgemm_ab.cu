// Timing harness for the non-transposed product D (M x N) = A (M x K) * B (K x N) with k_dgemm<false, true>
// (basis formation Psi = S (Omega Lambda^-1/2); reduced model X = Psi Z).   ./dgemm_ab M N K [nsplit]
#include <cstdio>
#include <vector>
#include "../../reducedbasismethods.jl_b200/csrc/dgemm_kernels.cuh"
using namespace rbm;
__global__ void fill(double *p, size_t n) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { unsigned long long z = (i + 1) * 0x9E3779B97F4A7C15ull; z ^= z >> 29; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 32; p[i] = (double)(z >> 11) / 9007199254740992.0 - 0.5; } }
int main(int argc, char **argv)
{
    const int64_t M = atoll(argv[1]), N = atoll(argv[2]), K = atoll(argv[3]);
    const int nsplit = argc > 4 ? atoi(argv[4]) : 1;
    double *A, *B, *D; cudaMalloc(&A, sizeof(double) * M * K); cudaMalloc(&B, sizeof(double) * K * N); cudaMalloc(&D, sizeof(double) * M * N * nsplit);
    fill<<<1024, 256>>>(A, (size_t)M * K); fill<<<1024, 256>>>(B, (size_t)K * N);
    std::vector<const double *> ha(K), hb(N);
    for (int64_t k = 0; k < K; ++k) ha[k] = A + k * M;
    for (int64_t n = 0; n < N; ++n) hb[n] = B + n * K;
    const double **da, **db; cudaMalloc(&da, sizeof(double *) * K); cudaMalloc(&db, sizeof(double *) * N);
    cudaMemcpy(da, ha.data(), sizeof(double *) * K, cudaMemcpyHostToDevice); cudaMemcpy(db, hb.data(), sizeof(double *) * N, cudaMemcpyHostToDevice);
    const int tm = (int)((M + GM_BM - 1) / GM_BM), tn = (int)((N + GM_BN - 1) / GM_BN);
    int64_t kps = (K + nsplit - 1) / nsplit; kps = (kps + GM_BK - 1) / GM_BK * GM_BK;
    GemmArgs a{}; a.acol = da; a.bcol = db; a.M = M; a.N = N; a.K = K; a.k_per_split = kps; a.out = D; a.ldo = M; a.split_stride = M * N; a.tiles_m = tm; a.same_ab = 0;
    if (argc > 6 && atoi(argv[6])) { a.a_base = A; a.a_ld = M; }
    const size_t smem = sizeof(double) * GM_STAGES * (GM_A_ELEMS_N + GM_B_ELEMS) + sizeof(double *) * (GM_BM + GM_BN);
    cudaFuncSetAttribute(k_dgemm<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_dgemm_pmb<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gm_pmb_smem<false>());
    const int mode = argc > 5 ? atoi(argv[5]) : 0;   // 1: persistent mbarrier kernel, compared with k_dgemm
    double *D2; cudaMalloc(&D2, sizeof(double) * M * N * nsplit); cudaMemset(D2, 0, sizeof(double) * M * N * nsplit);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dim3 grid(tm * tn, (unsigned)((K + kps - 1) / kps));
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (mode) { const int nitems = (int)(grid.x * grid.y); k_dgemm_pmb<false><<<nitems < sms ? nitems : sms, GM_THREADS, gm_pmb_smem<false>()>>>(a, (int)grid.x, nitems); }
        else k_dgemm<false, true><<<grid, GM_THREADS, smem>>>(a);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("M=%lld N=%lld K=%lld grid %u x %u: %.1f us  %.2f TF (%s)\n", (long long)M, (long long)N, (long long)K, grid.x, grid.y, ms * 1e3, 2.0 * M * N * K / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    if (mode) {
        GemmArgs a2 = a; a2.out = D2;
        k_dgemm<false, true><<<grid, GM_THREADS, smem>>>(a2);
        std::vector<double> h1((size_t)nsplit * M * N), h2(h1.size());
        cudaMemcpy(h1.data(), D, sizeof(double) * h1.size(), cudaMemcpyDeviceToHost);
        cudaMemcpy(h2.data(), D2, sizeof(double) * h2.size(), cudaMemcpyDeviceToHost);
        double md = 0, mx = 0;
        for (size_t q = 0; q < h1.size(); ++q) { double d1 = h1[q] - h2[q]; if (d1 < 0) d1 = -d1; if (d1 > md) md = d1; double v = h2[q] < 0 ? -h2[q] : h2[q]; if (v > mx) mx = v; }
        printf("pmb vs k_dgemm: max |diff| %.3e (max |value| %.3e) %s\n", md, mx, cudaGetErrorString(cudaDeviceSynchronize()));
    }
    return 0;
}
