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
    const double *a_base;       // A B only, optional: A is ONE column-major matrix, column k at a_base + k * a_ld
    int64_t a_ld;               //   (saves the per-slab loads of the column pointers); NULL: use acol
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
    // not volatile: a pure function of its register operands, so the scheduler may interleave it with the cp.async issue
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// One quarter (part = 0..3) of a [ncols=128][GM_BK] slab: column c of the operand, rows k0..k0+15, into
// smem[c*GM_LDK + k].  ptab: the tile's 128 column base pointers in shared memory (nullptr past the last column).
// VEC: 16-byte cp.async (needs even leading dimensions/aligned bases).
template <bool VEC>
__device__ __forceinline__ void load_part_colk(double *sm, const double *const *ptab, const void *dummy, int64_t k0, int64_t kend,
                                               int tid, int part)
{
    if (VEC) {
        constexpr int CH = GM_BK / 2;                       // chunks of 2 doubles per column
        constexpr int PER = GM_BN * CH / GM_THREADS / 4;    // chunks per thread per part
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int q = tid + GM_THREADS * (part * PER + u);
            const int c = q / CH, ch = q % CH;
            const int64_t k = k0 + ch * 2;
            const double *base = ptab[c];
            const bool ok = base != nullptr && k < kend;
            const double *src = ok ? base + k : nullptr;
            double *dst = sm + c * GM_LDK + ch * 2;
            if (ok && k + 1 >= kend) {  // odd tail: a single valid row
                cp_async_8(dst, src, true);
                dst[1] = 0.0;
            } else {
                cp_async_16(dst, ok ? (const void *)src : dummy, ok);  // !ok: zero-fill (dummy: any valid global address)
            }
        }
    } else {
        constexpr int PER = GM_BN * GM_BK / GM_THREADS / 4;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int q = tid + GM_THREADS * (part * PER + u);
            const int c = q / GM_BK, kk = q % GM_BK;
            const int64_t k = k0 + kk;
            const double *base = ptab[c];
            const bool ok = base != nullptr && k < kend;
            cp_async_8(sm + c * GM_LDK + kk, ok ? (const void *)(base + k) : dummy, ok);
        }
    }
}

// One quarter of a [GM_BK][128] slab of a non-transposed A: column k of A, rows m0..m0+127, into smem[kk*GM_LDM + m].
template <bool VEC>
__device__ __forceinline__ void load_part_km(double *sm, const double *const *cols, int64_t m0, int64_t M, int64_t k0,
                                             int64_t kend, int tid, int part)
{
    if (VEC) {
        constexpr int PER = GM_BK * (GM_BM / 2) / GM_THREADS / 4;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int q = tid + GM_THREADS * (part * PER + u);
            const int kk = q >> 6, ch = q & 63;
            const int64_t k = k0 + kk, m = m0 + ch * 2;
            const bool ok = k < kend && m < M;
            const double *src = ok ? cols[k] + m : cols[0];
            double *dst = sm + kk * GM_LDM + ch * 2;
            if (ok && m + 1 >= M) {
                cp_async_8(dst, src, true);
                dst[1] = 0.0;
            } else {
                cp_async_16(dst, src, ok);
            }
        }
    } else {
        constexpr int PER = GM_BK * GM_BM / GM_THREADS / 4;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int q = tid + GM_THREADS * (part * PER + u);
            const int kk = q >> 7, mm = q & 127;
            const int64_t k = k0 + kk, m = m0 + mm;
            const bool ok = k < kend && m < M;
            cp_async_8(sm + kk * GM_LDM + mm, ok ? cols[k] + m : cols[0], ok);
        }
    }
}

// Pipeline: GM_STAGES shared-memory buffers, slab s lives in buffer s % GM_STAGES.  Every fragment is read into registers
// one k4-step ahead (double-buffered af/bf), the loads of slab s+3 are issued in four quarters between the DMMA groups
// of slab s, and the only CTA barrier per slab sits before the LAST DMMA group of the slab: when a warp leaves it, it
// still has 32 DMMAs to issue, which hide the latency of the first fragment loads of the next slab.
template <bool TRANS_A, bool VEC>
__global__ void __launch_bounds__(GM_THREADS, 1) k_dgemm(const __grid_constant__ GemmArgs a)
{
    extern __shared__ double smem[];
    constexpr int A_ELEMS = TRANS_A ? GM_A_ELEMS_T : GM_A_ELEMS_N;
    constexpr int STAGE = A_ELEMS + GM_B_ELEMS;
    static_assert(GM_BK == 16 && GM_THREADS == 256 && GM_BM == 128 && GM_BN == 128, "load partition assumes this shape");
    const double **ptab = reinterpret_cast<const double **>(smem + GM_STAGES * STAGE);   // [128] A columns, [128] B columns
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

    if (tid < GM_BM) {
        const int64_t c = m0 + tid;
        ptab[tid] = (TRANS_A && c < a.M) ? a.acol[c] : nullptr;
    } else {
        const int64_t c = n0 + (tid - GM_BM);
        ptab[tid] = c < a.N ? a.bcol[c] : nullptr;
    }
    __syncthreads();

    double acc[GM_MI][GM_NJ][2];
#pragma unroll
    for (int i = 0; i < GM_MI; ++i)
#pragma unroll
        for (int j = 0; j < GM_NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int64_t nslab = kend > kbeg ? (kend - kbeg + GM_BK - 1) / GM_BK : 0;
    auto issue_part = [&](int64_t slab, int part) {
#ifdef GM_EXP_NOLOAD
        return;
#endif
        if (slab >= nslab) return;
        double *sa = smem + (slab % GM_STAGES) * STAGE;
        double *sb = sa + A_ELEMS;
        const int64_t k0 = kbeg + slab * GM_BK;
        if (TRANS_A) load_part_colk<VEC>(sa, ptab, a.bcol, k0, kend, tid, part);
        else load_part_km<VEC>(sa, a.acol, m0, a.M, k0, kend, tid, part);
        if (!diag) load_part_colk<VEC>(sb, ptab + GM_BM, a.bcol, k0, kend, tid, part);
    };
    auto load_frags = [&](int64_t slab, int k4, double (&af)[GM_MI], double (&bf)[GM_NJ]) {
        const double *sa = smem + (slab % GM_STAGES) * STAGE;
        const double *sb = diag ? sa : sa + A_ELEMS;
#pragma unroll
        for (int i = 0; i < GM_MI; ++i) {
            if (TRANS_A) af[i] = sa[(wm * (GM_MI * 8) + i * 8 + g) * GM_LDK + k4 * 4 + t4];
            else af[i] = sa[(k4 * 4 + t4) * GM_LDM + wm * (GM_MI * 8) + i * 8 + g];
        }
#pragma unroll
        for (int j = 0; j < GM_NJ; ++j) bf[j] = sb[(wn * (GM_NJ * 8) + j * 8 + g) * GM_LDK + k4 * 4 + t4];
    };
#pragma unroll
    for (int s = 0; s < GM_STAGES - 1; ++s) {
#pragma unroll
        for (int part = 0; part < 4; ++part) issue_part(s, part);
        cp_async_commit();
    }
    cp_async_wait<GM_STAGES - 2>();
    __syncthreads();
    double af[2][GM_MI], bf[2][GM_NJ];
    if (nslab > 0) load_frags(0, 0, af[0], bf[0]);
    for (int64_t slab = 0; slab < nslab; ++slab) {
#pragma unroll
        for (int k4 = 0; k4 < GM_BK / 4; ++k4) {
            // slab+3 goes into the buffer of slab-1, whose last fragment was read before the previous barrier
            issue_part(slab + GM_STAGES - 1, k4);
            if (k4 < GM_BK / 4 - 1) {
                load_frags(slab, k4 + 1, af[(k4 + 1) & 1], bf[(k4 + 1) & 1]);
            } else {
                cp_async_commit();
                cp_async_wait<GM_STAGES - 2>();   // slab+1 has landed (this thread's part) ...
#ifndef GM_EXP_NOBAR
                __syncthreads();                  // ... everybody's part; and nobody reads slab's buffer any more
#endif
                if (slab + 1 < nslab) load_frags(slab + 1, 0, af[0], bf[0]);
            }
#pragma unroll
            for (int i = 0; i < GM_MI; ++i)
#pragma unroll
                for (int j = 0; j < GM_NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[k4 & 1][i], bf[k4 & 1][j]);
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

// ---- mbarrier-pipelined variant of the A'B kernel (aligned columns) ------------------------------------------------
// Same tile, warp layout and padded shared-memory slabs as k_dgemm<true, true>; what changes is the pipeline control:
//   * every thread keeps the source pointers of its eight 16-byte chunks in registers and advances them by one slab per
//     iteration -- the steady state issues 8 LDGSTS and 8 pointer bumps per slab with no bounds logic (columns past the
//     edge of the matrix are clamped to the last valid one: they only feed accumulators that are never stored; a
//     trailing partial slab is staged separately with guarded loads);
//   * full[b]  (256 arrivals): cp.async.mbarrier.arrive.noinc -- slab b has landed; consumers wait on its parity;
//   * empty[b] (8 arrivals, one per warp after the DMMAs that consumed the slab's last fragments): producers wait on it
//     before refilling -- half a slab after the release, so the wait is normally already satisfied.
// The main loop therefore has no CTA-wide barrier and no cp.async.wait_group.
// (A cp.async.bulk version -- one 128-byte bulk copy per thread and slab -- was measured 17 % SLOWER: UBLKCP takes
// uniform registers, so per-thread bulk copies are serialised lane by lane.)
constexpr unsigned GM_SPIN_LIMIT = 1u << 26;   // bounded waits: a protocol bug traps instead of hanging the GPU

__device__ __forceinline__ void dmma884v(double &c0, double &c1, double a, double b)
{
    // volatile: program order against the mbarrier operations is part of the protocol (release AFTER the consuming DMMAs)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_on_cp_async(uint64_t *bar)
{
    // arrives (without raising the pending count) once all of this thread's earlier cp.async have completed
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    const unsigned addr = smem_u32(bar);
    for (unsigned spin = 0;; ++spin) {
        unsigned done;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done)
                     : "r"(addr), "r"(parity)
                     : "memory");
        if (done) return;
        if (spin > GM_SPIN_LIMIT) __trap();
    }
}
__device__ __forceinline__ void cp_async_16_full(unsigned dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
}

constexpr size_t GM_MB_SMEM = sizeof(double) * GM_STAGES * (GM_A_ELEMS_T + GM_B_ELEMS) + sizeof(uint64_t) * 2 * GM_STAGES;

#ifndef RBM_NO_ENTRY_POINTS   // a second translation unit (recon.cu) includes this header for its device functions only
__global__ void __launch_bounds__(GM_THREADS, 1) k_dgemm_mb(const __grid_constant__ GemmArgs a)
{
    extern __shared__ double smem[];
    constexpr int STAGE = GM_A_ELEMS_T + GM_B_ELEMS;
    constexpr int NK4 = GM_BK / 4;
    constexpr int CH = GM_BK / 2;                     // 16-byte chunks per column and slab
    constexpr int NCH = GM_BM * CH / GM_THREADS;      // chunks per thread, operand and slab: 4
    static_assert(NCH == 4 && NK4 == 4 && GM_BM == GM_BN, "issue schedule assumes 4 chunks per operand");
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + GM_STAGES * STAGE);
    uint64_t *empty = full + GM_STAGES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % GM_WM, wn = warp / GM_WM;
    const int g = lane >> 2, t4 = lane & 3;
    int2 tile;
    {
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
    const int64_t kbeg = (int64_t)blockIdx.y * a.k_per_split;
    int64_t kend = kbeg + a.k_per_split;
    if (kend > a.K) kend = a.K;
    const bool diag = a.same_ab && tile.x == tile.y;

    // this thread's chunks: column tid/8 + 32u of the tile, rows (tid%8)*2, +1 of the slab
    const int ccol = tid / CH, cch = tid % CH;
    const double *pa[NCH], *pb[NCH];
#pragma unroll
    for (int u = 0; u < NCH; ++u) {
        int64_t ca = m0 + ccol + 32 * u, cb = n0 + ccol + 32 * u;
        if (ca >= a.M) ca = a.M - 1;   // clamped duplicates feed only accumulators that are never stored
        if (cb >= a.N) cb = a.N - 1;
        pa[u] = a.acol[ca] + kbeg + cch * 2;
        pb[u] = a.bcol[cb] + kbeg + cch * 2;
    }
    const unsigned dst0 = smem_u32(smem) + (unsigned)((ccol * GM_LDK + cch * 2) * sizeof(double));
    if (tid == 0) {
        for (int b = 0; b < GM_STAGES; ++b) {
            mbar_init(full + b, GM_THREADS);
            mbar_init(empty + b, GM_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    double acc[GM_MI][GM_NJ][2];
#pragma unroll
    for (int i = 0; i < GM_MI; ++i)
#pragma unroll
        for (int j = 0; j < GM_NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int64_t nfull = kend > kbeg ? (kend - kbeg) / GM_BK : 0;   // whole slabs: pipelined
    const int ktail = kend > kbeg ? (int)((kend - kbeg) % GM_BK) : 0;
    // chunks u0 .. u0+1 of the next slab to stage (the pointers always address it) into buffer b
    auto issue_half = [&](int b, int u0) {
        const unsigned d = dst0 + (unsigned)(b * STAGE * sizeof(double));
#pragma unroll
        for (int u = u0; u < u0 + 2; ++u) {
            cp_async_16_full(d + (unsigned)(32 * u * GM_LDK * sizeof(double)), pa[u]);
            pa[u] += GM_BK;
            if (!diag) {
                cp_async_16_full(d + (unsigned)((GM_A_ELEMS_T + 32 * u * GM_LDK) * sizeof(double)), pb[u]);
                pb[u] += GM_BK;
            }
        }
    };
    auto load_frags = [&](int b, int k4, double (&af)[GM_MI], double (&bf)[GM_NJ]) {
        const double *sa = smem + b * STAGE;
        const double *sb = diag ? sa : sa + GM_A_ELEMS_T;
#pragma unroll
        for (int i = 0; i < GM_MI; ++i) af[i] = sa[(wm * (GM_MI * 8) + i * 8 + g) * GM_LDK + k4 * 4 + t4];
#pragma unroll
        for (int j = 0; j < GM_NJ; ++j) bf[j] = sb[(wn * (GM_NJ * 8) + j * 8 + g) * GM_LDK + k4 * 4 + t4];
    };
    auto mma_group = [&](const double (&af)[GM_MI], const double (&bf)[GM_NJ]) {
#pragma unroll
        for (int i = 0; i < GM_MI; ++i)
#pragma unroll
            for (int j = 0; j < GM_NJ; ++j) dmma884v(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    };

    double af[2][GM_MI], bf[2][GM_NJ];
    if (nfull > 0) {
#pragma unroll
        for (int s = 0; s < GM_STAGES - 1; ++s)
            if (s < nfull) {
                issue_half(s, 0);
                issue_half(s, 2);
                mbar_arrive_on_cp_async(full + s);
            }
        mbar_wait(full + 0, 0);
        load_frags(0, 0, af[0], bf[0]);
    }
    for (int64_t slab = 0; slab < nfull; ++slab) {
        const int b = (int)(slab % GM_STAGES);
        const int64_t nxt = slab + GM_STAGES - 1;            // staged during this iteration, into the buffer of slab-1
        const int nb = (int)(nxt % GM_STAGES);
        const bool refill = nxt < nfull;
#pragma unroll
        for (int k4 = 0; k4 < NK4; ++k4) {
            if (refill && k4 == NK4 / 2) {
                if (slab >= 1) mbar_wait(empty + nb, (unsigned)((slab - 1) / GM_STAGES) & 1u);   // half a slab of slack
                issue_half(nb, 0);
            }
            if (refill && k4 == NK4 / 2 + 1) {
                issue_half(nb, 2);
                mbar_arrive_on_cp_async(full + nb);
            }
            if (k4 < NK4 - 1) {
                load_frags(b, k4 + 1, af[(k4 + 1) & 1], bf[(k4 + 1) & 1]);
            } else if (slab + 1 < nfull) {
                mbar_wait(full + (int)((slab + 1) % GM_STAGES), (unsigned)((slab + 1) / GM_STAGES) & 1u);
                load_frags((int)((slab + 1) % GM_STAGES), 0, af[0], bf[0]);
            }
            mma_group(af[k4 & 1], bf[k4 & 1]);
        }
        // the DMMAs above consumed the last fragments read from buffer b: release it
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + b);
    }
    if (ktail > 0) {
        // trailing partial slab: plain guarded loads into buffer 0, zero-padded to a whole slab (pa/pb address it)
        __syncthreads();
#pragma unroll
        for (int u = 0; u < NCH; ++u) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const bool ok = cch * 2 + e < ktail;
                double *da = smem + (ccol + 32 * u) * GM_LDK + cch * 2 + e;
                da[0] = ok ? pa[u][e] : 0.0;
                if (!diag) da[GM_A_ELEMS_T] = ok ? pb[u][e] : 0.0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k4 = 0; k4 < NK4; ++k4) {
            load_frags(0, k4, af[0], bf[0]);
            mma_group(af[0], bf[0]);
        }
    }
    double *out = a.out + (size_t)blockIdx.y * a.split_stride;
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
#endif

// ---- persistent mbarrier-pipelined kernel (aligned columns), both A'B and A B ----------------------------------------
// k_dgemm_mb above runs one (tile, split) item per CTA: pipeline fill and the epilogue of every item are exposed, which
// costs ~40 % on short contractions (X = Psi Z of the reduced model: 11 slabs per item).  Here one CTA per SM walks its
// items w = blockIdx.x, blockIdx.x + gridDim.x, ... and the slab pipeline runs across item boundaries: while the last
// slabs of item i are multiplied (and while its accumulators are stored) the first slabs of item i+1 are already in
// flight.  Producer and consumer keep separate cursors (item, slab within the item, global slab number = buffer/parity);
// the producer stays exactly GM_STAGES-1 slabs ahead.  Edges and the partial last slab of an item go through the same
// pipeline with the zero-filling form of cp.async (src-size 0, 8 or 16).
struct GemmItem {
    int64_t m0, n0, kbeg;
    int nslab, nfull, ktail, split;
    bool diag;
};

__device__ __forceinline__ void cp_async_16_sz(unsigned dst, const void *src, unsigned bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

template <bool TRANS_A>
__global__ void __launch_bounds__(GM_THREADS, 1) k_dgemm_pmb(const __grid_constant__ GemmArgs a, int ntiles, int nitems)
{
    extern __shared__ double smem[];
    constexpr int A_ELEMS = TRANS_A ? GM_A_ELEMS_T : GM_A_ELEMS_N;
    constexpr int STAGE = A_ELEMS + GM_B_ELEMS;
    constexpr int NK4 = GM_BK / 4;
    constexpr int CH = GM_BK / 2;
    static_assert(GM_BM * CH / GM_THREADS == 4 && GM_BK * (GM_BM / 2) / GM_THREADS == 4 && NK4 == 4 && GM_BM == GM_BN,
                  "issue schedule assumes 4 chunks per thread, operand and slab");
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + GM_STAGES * STAGE);
    uint64_t *empty = full + GM_STAGES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % GM_WM, wn = warp / GM_WM;
    const int g = lane >> 2, t4 = lane & 3;
    const unsigned smem0 = smem_u32(smem);

    auto item_info = [&](int w) {
        GemmItem it;
        int t = w % ntiles, j = 0;
        it.split = w / ntiles;
        int tx, ty;
        if (a.same_ab) {
            while (t >= a.tiles_m - j) {
                t -= a.tiles_m - j;
                ++j;
            }
            tx = j + t; ty = j;
        } else {
            tx = t % a.tiles_m; ty = t / a.tiles_m;
        }
        it.m0 = (int64_t)tx * GM_BM;
        it.n0 = (int64_t)ty * GM_BN;
        it.kbeg = (int64_t)it.split * a.k_per_split;
        int64_t kend = it.kbeg + a.k_per_split;
        if (kend > a.K) kend = a.K;
        const int64_t len = kend > it.kbeg ? kend - it.kbeg : 0;
        it.nfull = (int)(len / GM_BK);
        it.ktail = (int)(len % GM_BK);
        it.nslab = it.nfull + (it.ktail ? 1 : 0);
        it.diag = a.same_ab && tx == ty;
        return it;
    };

    if (tid == 0) {
        for (int b = 0; b < GM_STAGES; ++b) {
            mbar_init(full + b, GM_THREADS);
            mbar_init(empty + b, GM_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    // ---------------------------------------------------------------- producer (every thread stages its 8 chunks)
    // [col][k] slabs (B always, A when TRANS_A): column tid/8 + 32u of the tile, rows (tid%8)*2, +1 of the slab
    // [k][m] slab (A when !TRANS_A):            row tid/64 + 4u of the slab, tile rows (tid%64)*2, +1
    const int ccol = tid / CH, cch = tid % CH;
    const int akk = tid >> 6, ach = tid & 63;
    int pw = blockIdx.x;          // producer's item
    int ps = 0;                   // slab within the item
    unsigned pg = 0;              // global slab number of this CTA (buffer = pg % GM_STAGES)
    bool pactive = pw < nitems;
    GemmItem pi{};
    const double *pa[4], *pb[4];  // TRANS_A: running chunk pointers; !TRANS_A: pa = this slab's four A-column pointers
    unsigned sz_m = 16;           // !TRANS_A: bytes of this thread's A chunks that lie inside the matrix (rows m, m+1)
    int64_t arow = 0;             // !TRANS_A: first of this thread's two rows of A
    auto load_a_cols = [&]() {    // !TRANS_A: column pointers of slab ps (clamped past the end of the contraction)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int64_t k = pi.kbeg + (int64_t)ps * GM_BK + akk + 4 * u;
            if (k >= a.K) k = a.K - 1;
            pa[u] = (a.a_base ? a.a_base + k * a.a_ld : a.acol[k]) + arow;
        }
    };
    auto setup_producer = [&]() {
        pi = item_info(pw);
        ps = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int64_t cb = pi.n0 + ccol + 32 * u;
            if (cb >= a.N) cb = a.N - 1;   // clamped duplicates feed only accumulators that are never stored
            pb[u] = a.bcol[cb] + pi.kbeg + cch * 2;
            if (TRANS_A) {
                int64_t ca = pi.m0 + ccol + 32 * u;
                if (ca >= a.M) ca = a.M - 1;
                pa[u] = a.acol[ca] + pi.kbeg + cch * 2;
            }
        }
        if (!TRANS_A) {
            const int64_t left = a.M - (pi.m0 + ach * 2);
            sz_m = left >= 2 ? 16u : (left == 1 ? 8u : 0u);
            arow = left > 0 ? pi.m0 + ach * 2 : 0;                      // never dereferenced when sz_m == 0, but kept valid
            load_a_cols();
        }
    };
    // chunks 2h, 2h+1 of producer slab (pi, ps) into buffer pg % GM_STAGES
    auto issue_half = [&](int h) {
        const unsigned base = smem0 + (unsigned)((pg % GM_STAGES) * STAGE * sizeof(double));
        const int rows = ps < pi.nfull ? GM_BK : pi.ktail;                 // valid rows of this slab
        const int rk = rows - cch * 2;
        const unsigned sz_k = rk >= 2 ? 16u : (rk == 1 ? 8u : 0u);         // [col][k] chunks: rows cch*2, +1
#pragma unroll
        for (int u = 2 * h; u < 2 * h + 2; ++u) {
            if (TRANS_A) {
                cp_async_16_sz(base + (unsigned)(((ccol + 32 * u) * GM_LDK + cch * 2) * sizeof(double)), pa[u], sz_k);
                pa[u] += GM_BK;
            } else {
                cp_async_16_sz(base + (unsigned)(((akk + 4 * u) * GM_LDM + ach * 2) * sizeof(double)), pa[u],
                               akk + 4 * u < rows ? sz_m : 0u);
            }
            if (!pi.diag) {
                cp_async_16_sz(base + (unsigned)((A_ELEMS + (ccol + 32 * u) * GM_LDK + cch * 2) * sizeof(double)), pb[u], sz_k);
                pb[u] += GM_BK;
            }
        }
    };
    auto finish_slab = [&]() {
        mbar_arrive_on_cp_async(full + pg % GM_STAGES);
        ++pg;
        ++ps;
        if (ps == pi.nslab) {
            pw += gridDim.x;
            pactive = pw < nitems;
            if (pactive) setup_producer();
        } else if (!TRANS_A) {
            load_a_cols();
        }
    };

    // ---------------------------------------------------------------- consumer
    auto load_frags = [&](int b, bool diag, int k4, double (&af)[GM_MI], double (&bf)[GM_NJ]) {
        const double *sa = smem + b * STAGE;
        const double *sb = diag ? sa : sa + A_ELEMS;
#pragma unroll
        for (int i = 0; i < GM_MI; ++i) {
            if (TRANS_A) af[i] = sa[(wm * (GM_MI * 8) + i * 8 + g) * GM_LDK + k4 * 4 + t4];
            else af[i] = sa[(k4 * 4 + t4) * GM_LDM + wm * (GM_MI * 8) + i * 8 + g];
        }
#pragma unroll
        for (int j = 0; j < GM_NJ; ++j) bf[j] = sb[(wn * (GM_NJ * 8) + j * 8 + g) * GM_LDK + k4 * 4 + t4];
    };
    double acc[GM_MI][GM_NJ][2];
    auto mma_group = [&](const double (&af)[GM_MI], const double (&bf)[GM_NJ]) {
#pragma unroll
        for (int i = 0; i < GM_MI; ++i)
#pragma unroll
            for (int j = 0; j < GM_NJ; ++j) {
                // !TRANS_A: accumulate D' (operands swapped; the m8n8k4 fragments of A and B have the same lane
                // layout), so that a thread owns rows 2*t4, 2*t4+1 of column g: 16-byte stores into column-major D
                if (TRANS_A) dmma884v(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                else dmma884v(acc[i][j][0], acc[i][j][1], bf[j], af[i]);
            }
    };

    if (pactive) setup_producer();
#pragma unroll 1
    for (int s = 0; s < GM_STAGES - 1 && pactive; ++s) {   // pipeline fill (may already run into the second item)
        issue_half(0);
        issue_half(1);
        finish_slab();
    }
    unsigned cg = 0;
    double af[2][GM_MI], bf[2][GM_NJ];
    for (int cw = blockIdx.x; cw < nitems; cw += gridDim.x) {
        const GemmItem ci = item_info(cw);
#pragma unroll
        for (int i = 0; i < GM_MI; ++i)
#pragma unroll
            for (int j = 0; j < GM_NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        if (ci.nslab > 0) {
            mbar_wait(full + cg % GM_STAGES, (cg / GM_STAGES) & 1u);
            load_frags((int)(cg % GM_STAGES), ci.diag, 0, af[0], bf[0]);
        }
#pragma unroll 1
        for (int s = 0; s < ci.nslab; ++s, ++cg) {
            const int b = (int)(cg % GM_STAGES);
#pragma unroll
            for (int k4 = 0; k4 < NK4; ++k4) {
                if (k4 == NK4 / 2 && pactive) {
                    // refill the buffer released GM_STAGES slabs ago (half a slab of slack since its release)
                    if (pg >= GM_STAGES) mbar_wait(empty + pg % GM_STAGES, (pg / GM_STAGES - 1) & 1u);
                    issue_half(0);
                }
                if (k4 == NK4 / 2 + 1 && pactive) {
                    issue_half(1);
                    finish_slab();
                }
                if (k4 < NK4 - 1) {
                    load_frags(b, ci.diag, k4 + 1, af[(k4 + 1) & 1], bf[(k4 + 1) & 1]);
                } else if (s + 1 < ci.nslab) {
                    mbar_wait(full + (cg + 1) % GM_STAGES, ((cg + 1) / GM_STAGES) & 1u);
                    load_frags((int)((cg + 1) % GM_STAGES), ci.diag, 0, af[0], bf[0]);
                }
                mma_group(af[k4 & 1], bf[k4 & 1]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + b);   // after the DMMAs that consumed the slab's last fragments
        }
        double *out = a.out + (size_t)ci.split * a.split_stride;
        if (TRANS_A) {
#pragma unroll
            for (int i = 0; i < GM_MI; ++i) {
                const int64_t row = ci.m0 + wm * (GM_MI * 8) + i * 8 + g;
                if (row >= a.M) continue;
#pragma unroll
                for (int j = 0; j < GM_NJ; ++j) {
                    const int64_t col = ci.n0 + wn * (GM_NJ * 8) + j * 8 + t4 * 2;
                    if (col < a.N) out[col * a.ldo + row] = acc[i][j][0];
                    if (col + 1 < a.N) out[(col + 1) * a.ldo + row] = acc[i][j][1];
                }
            }
        } else {
            const bool vec_out = (a.ldo & 1) == 0 && ((uintptr_t)out & 15) == 0;
#pragma unroll
            for (int j = 0; j < GM_NJ; ++j) {
                const int64_t col = ci.n0 + wn * (GM_NJ * 8) + j * 8 + g;
                if (col >= a.N) continue;
#pragma unroll
                for (int i = 0; i < GM_MI; ++i) {
                    const int64_t row = ci.m0 + wm * (GM_MI * 8) + i * 8 + t4 * 2;
                    double *dst = out + col * a.ldo + row;
                    if (vec_out && row + 1 < a.M) {
                        *reinterpret_cast<double2 *>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
                    } else {
                        if (row < a.M) dst[0] = acc[i][j][0];
                        if (row + 1 < a.M) dst[1] = acc[i][j][1];
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
}

template <bool TRANS_A> constexpr size_t gm_pmb_smem()
{
    return sizeof(double) * GM_STAGES * ((TRANS_A ? GM_A_ELEMS_T : GM_A_ELEMS_N) + GM_B_ELEMS) + sizeof(uint64_t) * 2 * GM_STAGES;
}

// out[i][j] = sum_s part[s][i][j]; symmetric: lower triangle summed, mirrored to the upper.
#ifndef RBM_NO_ENTRY_POINTS   // a second translation unit (recon.cu) includes this header for its device functions only
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
#endif

}  // namespace rbm
