// dgemm_kernels.cuh -- FP64 tensor-core (DMMA) tile kernel behind the snapshot POD.
//
// sm_100a has no FP64 kind of tcgen05.mma; FP64 tensor math is `mma.sync.m8n8k4.f64`
// (SASS DMMA.8x8x4 -- the only native FP64 shape: m16n8k16 decomposes into eight of
// them).  One kernel template serves the three dense contractions of the reference:
//   TRANS_A = true : D = A' B, contraction over the (long) row index
//        Gram      G = [X V]'[X V]      src/algorithms/evd.jl:32-33, S'S :65
//        project   Z = Psi' S           scripts/bump_on_tail_2_projections.jl:44-60
//   TRANS_A = false: D = A B, contraction over columns
//        basis     Psi = S (Omega_k |Lambda|^-1/2)   src/algorithms/evd.jl:49,81
// Operands are described by arrays of column pointers so that [X V] (and any list of
// column blocks) is never concatenated.
//
// CTA tile 128 x 128, K-slab 16, 4-stage cp.async pipeline (160 KB), 8 warps (2 x 4), warp tile
// 64 x 32 = 8 x 4 DMMA fragments (64 FP64 accumulators = 128 registers per thread).
// Shared-memory rows are padded to a stride == 4 (mod 16) doubles, which makes every
// fragment load (8 groups x 4 k) hit 16 distinct 64-bit banks per half-warp.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace rbm {

constexpr int GM_BM = 128, GM_BN = 128, GM_BK = 16, GM_STAGES = 4;
constexpr int GM_WM = 2, GM_WN = 4;                 // warps along M and N (4 x 4 measured 6 % slower)
constexpr int GM_THREADS = 32 * GM_WM * GM_WN;      // 256
constexpr int GM_MI = GM_BM / (GM_WM * 8), GM_NJ = GM_BN / (GM_WN * 8);   // 8x8 DMMA fragments per warp: 8 x 4
constexpr int GM_LDK = GM_BK + 4;   // [col][k] tiles (k contiguous)
constexpr int GM_LDM = GM_BM + 4;   // [k][m] tile of a non-transposed A (m contiguous)
constexpr int GM_A_ELEMS_T = GM_BM * GM_LDK;
constexpr int GM_A_ELEMS_N = GM_BK * GM_LDM;
constexpr int GM_B_ELEMS = GM_BN * GM_LDK;

struct GemmArgs {
    const double *const *acol;  // TRANS_A: M column pointers (length K each); else K column pointers (length M)
    const double *const *bcol;  // N column pointers (length K each)
    int64_t M, N, K;
    int64_t k_per_split;        // multiple of GM_BK
    double *out;                // [nsplit] matrices, column-major M x N with leading dimension ldo
    int64_t ldo, split_stride;
    int tiles_m;                // number of tile rows; blockIdx.x enumerates tiles column by column (sym: i >= j only)
    int same_ab;                // Gram: A and B are the same matrix (diagonal tiles load once)
};

__device__ __forceinline__ void cp_async_16(void *smem, const void *gmem, bool pred)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_8(void *smem, const void *gmem, bool pred)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Load a [ncols=128][GM_BK] slab: column c of the operand, rows k0..k0+15, into
// smem[c*GM_LDK + k] (rows k0..k0+GM_BK-1).  VEC: 16-byte cp.async (needs even leading dimensions/aligned bases).
template <bool VEC>
__device__ __forceinline__ void load_slab_colk(double *sm, const double *const *cols, int64_t col0, int64_t ncols_total,
                                               int64_t k0, int64_t kend, int tid)
{
    if (VEC) {
        // 128 cols x GM_BK/2 chunks of 2 doubles
        constexpr int CH = GM_BK / 2;
#pragma unroll
        for (int u = 0; u < GM_BN * CH / GM_THREADS; ++u) {
            int q = tid + GM_THREADS * u;
            int c = q / CH, part = q % CH;
            int64_t gc = col0 + c, k = k0 + part * 2;
            bool ok = gc < ncols_total && k < kend;
            const double *src = ok ? cols[gc] + k : cols[0];
            double *dst = sm + c * GM_LDK + part * 2;
            if (ok && k + 1 >= kend) {  // odd tail: a single valid row
                cp_async_8(dst, src, true);
                dst[1] = 0.0;
            } else {
                cp_async_16(dst, src, ok);  // !ok: zero-fill
            }
        }
    } else {
#pragma unroll
        for (int u = 0; u < GM_BN * GM_BK / GM_THREADS; ++u) {
            int q = tid + GM_THREADS * u;
            int c = q / GM_BK, kk = q % GM_BK;
            int64_t gc = col0 + c, k = k0 + kk;
            bool ok = gc < ncols_total && k < kend;
            cp_async_8(sm + c * GM_LDK + kk, ok ? cols[gc] + k : cols[0], ok);
        }
    }
}

// Load a [GM_BK][128] slab of a non-transposed A: column k of A, rows m0..m0+127, into
// smem[kk*GM_LDM + m].
template <bool VEC>
__device__ __forceinline__ void load_slab_km(double *sm, const double *const *cols, int64_t m0, int64_t M, int64_t k0,
                                             int64_t kend, int tid)
{
    if (VEC) {
#pragma unroll
        for (int u = 0; u < GM_BK * (GM_BM / 2) / GM_THREADS; ++u) {
            int q = tid + GM_THREADS * u;
            int kk = q >> 6, part = q & 63;
            int64_t k = k0 + kk, m = m0 + part * 2;
            bool ok = k < kend && m < M;
            const double *src = ok ? cols[k] + m : cols[0];
            double *dst = sm + kk * GM_LDM + part * 2;
            if (ok && m + 1 >= M) {
                cp_async_8(dst, src, true);
                dst[1] = 0.0;
            } else {
                cp_async_16(dst, src, ok);
            }
        }
    } else {
#pragma unroll
        for (int u = 0; u < GM_BK * GM_BM / GM_THREADS; ++u) {
            int q = tid + GM_THREADS * u;
            int kk = q >> 7, mm = q & 127;
            int64_t k = k0 + kk, m = m0 + mm;
            bool ok = k < kend && m < M;
            cp_async_8(sm + kk * GM_LDM + mm, ok ? cols[k] + m : cols[0], ok);
        }
    }
}

template <bool TRANS_A, bool VEC>
__global__ void __launch_bounds__(GM_THREADS, 1) k_dgemm(const __grid_constant__ GemmArgs a)
{
    extern __shared__ double smem[];
    constexpr int A_ELEMS = TRANS_A ? GM_A_ELEMS_T : GM_A_ELEMS_N;
    constexpr int STAGE = A_ELEMS + GM_B_ELEMS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % GM_WM, wn = warp / GM_WM;
    const int g = lane >> 2, t4 = lane & 3;
    int2 tile;
    {   // tile coordinates from the linear block index (no tile list in memory: the launch stays asynchronous)
        int t = blockIdx.x, j = 0;
        if (a.same_ab) {
            while (t >= a.tiles_m - j) {
                t -= a.tiles_m - j;
                ++j;
            }
            tile = make_int2(j + t, j);
        } else {
            tile = make_int2(t % a.tiles_m, t / a.tiles_m);
        }
    }
    const int64_t m0 = (int64_t)tile.x * GM_BM, n0 = (int64_t)tile.y * GM_BN;
    const int split = blockIdx.y;
    const int64_t kbeg = (int64_t)split * a.k_per_split;
    int64_t kend = kbeg + a.k_per_split;
    if (kend > a.K) kend = a.K;
    const bool diag = a.same_ab && tile.x == tile.y;

    double acc[GM_MI][GM_NJ][2];
#pragma unroll
    for (int i = 0; i < GM_MI; ++i)
#pragma unroll
        for (int j = 0; j < GM_NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int64_t nslab = kend > kbeg ? (kend - kbeg + GM_BK - 1) / GM_BK : 0;
    auto issue = [&](int64_t slab) {
        double *sa = smem + (slab % GM_STAGES) * STAGE;
        double *sb = sa + A_ELEMS;
        const int64_t k0 = kbeg + slab * GM_BK;
        if (TRANS_A) load_slab_colk<VEC>(sa, a.acol, m0, a.M, k0, kend, tid);
        else load_slab_km<VEC>(sa, a.acol, m0, a.M, k0, kend, tid);
        if (!diag) load_slab_colk<VEC>(sb, a.bcol, n0, a.N, k0, kend, tid);
    };
#pragma unroll
    for (int s = 0; s < GM_STAGES - 1; ++s) {
        if (s < nslab) issue(s);
        cp_async_commit();
    }
    for (int64_t slab = 0; slab < nslab; ++slab) {
        cp_async_wait<GM_STAGES - 2>();
        __syncthreads();
        if (slab + GM_STAGES - 1 < nslab) issue(slab + GM_STAGES - 1);
        cp_async_commit();
        const double *sa = smem + (slab % GM_STAGES) * STAGE;
        const double *sb = diag ? sa : sa + A_ELEMS;
#pragma unroll
        for (int k4 = 0; k4 < GM_BK / 4; ++k4) {
            double af[GM_MI], bf[GM_NJ];
#pragma unroll
            for (int i = 0; i < GM_MI; ++i) {
                if (TRANS_A) af[i] = sa[(wm * (GM_MI * 8) + i * 8 + g) * GM_LDK + k4 * 4 + t4];
                else af[i] = sa[(k4 * 4 + t4) * GM_LDM + wm * (GM_MI * 8) + i * 8 + g];
            }
#pragma unroll
            for (int j = 0; j < GM_NJ; ++j) bf[j] = sb[(wn * (GM_NJ * 8) + j * 8 + g) * GM_LDK + k4 * 4 + t4];
#pragma unroll
            for (int i = 0; i < GM_MI; ++i)
#pragma unroll
                for (int j = 0; j < GM_NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();
    // epilogue: C fragment (row g, cols 2*t4, 2*t4+1) -> column-major out
    double *out = a.out + (size_t)split * a.split_stride;
#pragma unroll
    for (int i = 0; i < GM_MI; ++i) {
        const int64_t row = m0 + wm * (GM_MI * 8) + i * 8 + g;
        if (row >= a.M) continue;
#pragma unroll
        for (int j = 0; j < GM_NJ; ++j) {
            const int64_t col = n0 + wn * (GM_NJ * 8) + j * 8 + t4 * 2;
            if (col < a.N) out[col * a.ldo + row] = acc[i][j][0];
            if (col + 1 < a.N) out[(col + 1) * a.ldo + row] = acc[i][j][1];
        }
    }
}

// out[i][j] = sum_s part[s][i][j]; symmetric: lower triangle summed, mirrored to the upper.
__global__ void k_splitk_reduce(const double *part, int nsplit, int64_t split_stride, int64_t M, int64_t N, int64_t ldp,
                                double *out, int64_t ldo, int sym)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * N) return;
    const int64_t i = idx % M, j = idx / M;
    if (sym && i < j) return;
    double s = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) s += part[(size_t)sp * split_stride + j * ldp + i];
    out[j * ldo + i] = s;
    if (sym && i != j) out[i * ldo + j] = s;
}

}  // namespace rbm
