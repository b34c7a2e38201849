// recon.cu -- reduced model: reconstruction X = Psi Z and charge deposition in ONE kernel.
//
// Reference: update!(f, Z, w, t) of the reduced field reconstructs all N particle positions of a sample,
// x = Psi * z, and deposits them (src/particles/electric_field.jl:22-25) -- DEIM hyper-reduces only the gather.
// As two kernels the product writes X (N x S doubles) to HBM and the deposit reads it back every time step, and
// neither hides the other: the product is bound by the FP64 tensor pipe, the deposit by instruction issue and the
// shared-memory pipe.  Here X exists only in the DMMA accumulators, the deposit is the epilogue of the product, and TWO
// CTAs share an SM so that the epilogue of one runs under the DMMA main loop of the other.
//
//   CTA = 4 warps (2 x 2), tile 128 rows x 64 samples, warp tile 64 x 32 = 8 x 4 DMMA.8x8x4 fragments (the fragment code
//   of dgemm_kernels.cuh: accumulation of D' so that a thread owns two consecutive rows of a sample), K-slab 16, 3-stage
//   cp.async pipeline (80 KB) under mbarrier control that runs across tile boundaries, up to 255 registers x 128 threads
//   x 2 CTAs = the register file.
//
// Accumulator layout: thread (g = lane/4, t4 = lane%4) of warp (wm, wn) holds, for i < 8, j < 4, e < 2, the entry of row
// m0 + 64 wm + 8 i + 2 t4 + e and sample n0 + 32 wn + 8 j + g: sixteen rows of FOUR samples.  A thread-private histogram
// column (plain LDS/DADD/STS, fixed summation order -- FP64 shared-memory atomics are CAS loops on sm_100a) must belong
// to ONE sample, so the four lanes of a quad (same g) exchange their values with three shuffles per (i, e): lane t4
// keeps sample j = t4 and receives that sample's values of the other three lanes.  Thread (g, t4) then owns sample
// n0 + 32 wn + 8 t4 + g and deposits its 64 rows of the warp tile into hist[bin][tid]; the two warps wm = 0, 1 that share
// a sample are added when the histogram is published.
//
// Rows past the end of the matrix are zero-filled by the loads, so their x is exactly 0: they deposit 6 B(0) = (1, 4, 1, 0)
// into bins 0..2, which the owning thread subtracts again (exact small integers) -- no per-particle mask.
//
// A CTA walks its tiles in order of increasing sample tile, publishes one partial per sample tile it has visited and
// zeros for the others: part[sample][cta][bin], summed over the CTAs in a fixed order by the field solve (k_solve),
// exactly like the chunk partials of the full-order deposit.  Bit-reproducible from run to run.
#include <cstdlib>

#include "common.cuh"
#define RBM_NO_ENTRY_POINTS   // device functions and kernel templates only: the plain kernels live in pic.cu / pod.cu
#include "pic_kernels.cuh"
#include "dgemm_kernels.cuh"

namespace rbm {

constexpr int RD_BN = 64, RD_BK = 16, RD_THREADS = 128;
constexpr int RD_NJ = 4;                                  // 8x8 fragments per warp along the samples: 32 samples
constexpr int RD_LDK = RD_BK + 4;                         // padded strides == 4 (mod 16) doubles: conflict-free fragment loads
constexpr int RD_B = RD_BN * RD_LDK;
// Shape variants: MI = 8x8 fragments per warp along the rows (warp tile 8 MI x 32, CTA tile 16 MI x 64), STAGES = depth of
// the operand pipeline, CTAS = CTAs per SM the register and shared-memory budgets are cut for.
template <int MI, int STAGES> struct ReconShape {
    static constexpr int BM = 16 * MI, LDM = BM + 4, A = RD_BK * LDM, STAGE = A + RD_B;
    static constexpr size_t smem(int nh)
    {
        return sizeof(double) * ((size_t)STAGES * STAGE + (size_t)(nh + 3) * RD_THREADS) + sizeof(uint64_t) * 2 * STAGES;
    }
};

struct ReconArgs {
    const double *A;   // N x K, column-major, lda even, 16-byte aligned (Psi)
    int64_t lda;
    const double *Z;   // K x S, column-major, ldz even, 16-byte aligned (reduced coordinates of the S samples)
    int64_t ldz;
    int64_t N;
    int S, K;
    int tiles_m, ntiles;
    Geom g;
    double *part;      // [S][gridDim.x][nh]
};

template <int RD_MI, int RD_STAGES, int CTAS>
__global__ void __launch_bounds__(RD_THREADS, CTAS) k_recon_deposit(const __grid_constant__ ReconArgs a)
{
    using Shape = ReconShape<RD_MI, RD_STAGES>;
    constexpr int RD_BM = Shape::BM, RD_LDM = Shape::LDM, RD_A = Shape::A, RD_STAGE = Shape::STAGE;
    constexpr int ACH = RD_BM / 2;                          // 16-byte chunks per slab row of A
    constexpr int NA = RD_BK * ACH / RD_THREADS;            // A chunks per thread and slab: 8 (MI = 8) or 4
    constexpr int AKS = RD_THREADS / ACH;                   // slab rows between a thread's consecutive A chunks: 2 or 4
    static_assert(NA % 2 == 0 && RD_BN * (RD_BK / 2) / RD_THREADS == 4, "issue schedule: two halves per slab");
    extern __shared__ double smem[];
    double *hist = smem + RD_STAGES * RD_STAGE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 1, wn = warp >> 1;
    const int g = lane >> 2, t4 = lane & 3;
    const int nh = a.g.nh;
    const int nslab = (a.K + RD_BK - 1) / RD_BK;
    const int my_tiles = (a.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_tiles * nslab;
    const unsigned smem0 = smem_u32(smem);

    auto zero_hist = [&]() {
        for (int q = tid; q < (nh + 3) * RD_THREADS; q += RD_THREADS) hist[q] = 0.0;
    };
    // partial of this CTA for the 64 samples of sample tile ty: from the histogram (have) or zeros
    auto publish = [&](int ty, bool have) {
        if (have) __syncthreads();
        for (int f = tid; f < RD_BN * nh; f += RD_THREADS) {
            const int sl = f / nh, b = f - sl * nh;
            const int s = ty * RD_BN + sl;
            if (s >= a.S) break;
            double v = 0.0;
            if (have) {
                // columns of sample sl = 32 wn + 8 t4 + g: thread 4 g + t4 of warps 2 wn (wm = 0) and 2 wn + 1 (wm = 1)
                const int c0 = (sl >> 5) * 64 + (sl & 7) * 4 + ((sl >> 3) & 3);
                const double *r0 = hist + b * RD_THREADS + c0;
                v = r0[0] + r0[32];
                if (b < 3) {   // alias rows nh..nh+2
                    const double *r1 = hist + (nh + b) * RD_THREADS + c0;
                    v += r1[0] + r1[32];
                }
            }
            a.part[((size_t)s * gridDim.x + blockIdx.x) * nh + b] = v;
        }
        if (have) {
            __syncthreads();
            zero_hist();
            __syncthreads();
        }
    };

    // ---------------------------------------------------------------- producer (every thread stages its 12 chunks per slab)
    // Pipeline control as in k_dgemm_pmb (dgemm_kernels.cuh): full[b] collects cp.async.mbarrier.arrive.noinc from the 128
    // threads, empty[b] one arrival per warp after the DMMAs that consumed the slab; the refill of a buffer is issued half
    // a slab after its release.  No CTA barrier and no cp.async.wait_group in the main loop.
    uint64_t *full = reinterpret_cast<uint64_t *>(hist + (nh + 3) * RD_THREADS);
    uint64_t *empty = full + RD_STAGES;
    if (tid == 0) {
        for (int b = 0; b < RD_STAGES; ++b) {
            mbar_init(full + b, RD_THREADS);
            mbar_init(empty + b, RD_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    zero_hist();
    __syncthreads();

    // A [k][m] slab: slab rows akk + AKS u (u < NA), tile rows 2 ach, +1;  Z [sample][k] slab: samples bcol + 16u (u < 4), rows 2 bch, +1
    const int akk = tid / ACH, ach = tid % ACH, bcol = tid >> 3, bch = tid & 7;
    unsigned pg = 0;               // flat slab number of this CTA (buffer = pg % RD_STAGES)
    int ps = 0, pt = blockIdx.x;   // slab within the tile, tile
    bool pactive = total > 0;
    const double *pa = a.A, *pb[4] = {a.Z, a.Z, a.Z, a.Z};
    unsigned sz_m = 0;             // bytes of this thread's A chunks that lie inside the matrix
    auto producer_tile = [&]() {
        const int64_t m = (int64_t)(pt % a.tiles_m) * RD_BM + ach * 2, left = a.N - m;
        const int n0 = (pt / a.tiles_m) * RD_BN;
        sz_m = left >= 2 ? 16u : (left == 1 ? 8u : 0u);
        pa = a.A + (left > 0 ? m : 0) + (int64_t)akk * a.lda;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int n = n0 + bcol + 16 * u;
            if (n >= a.S) n = a.S - 1;   // clamped duplicates feed only histogram columns that are never published
            pb[u] = a.Z + (int64_t)n * a.ldz + bch * 2;
        }
    };
    auto issue_half = [&](int h) {   // half of the chunks of producer slab (pt, ps) into buffer pg % RD_STAGES
        const unsigned base = smem0 + (unsigned)((pg % RD_STAGES) * RD_STAGE * sizeof(double));
        const int rows = a.K - ps * RD_BK;   // valid rows of this slab (>= RD_BK except in the last one)
        const int rk = rows - bch * 2;
        const unsigned sz_k = rk >= 2 ? 16u : (rk == 1 ? 8u : 0u);
#pragma unroll
        for (int u = (NA / 2) * h; u < (NA / 2) * (h + 1); ++u) {
            const bool in = akk + AKS * u < rows;
            cp_async_16_sz(base + (unsigned)(((akk + AKS * u) * RD_LDM + ach * 2) * sizeof(double)),
                           in ? pa + (int64_t)(AKS * u) * a.lda : a.A, in ? sz_m : 0u);
        }
#pragma unroll
        for (int u = 2 * h; u < 2 * h + 2; ++u)
            cp_async_16_sz(base + (unsigned)((RD_A + (bcol + 16 * u) * RD_LDK + bch * 2) * sizeof(double)), sz_k ? pb[u] : a.Z, sz_k);
    };
    auto finish_slab = [&]() {
        mbar_arrive_on_cp_async(full + pg % RD_STAGES);
        ++pg;
        if (++ps == nslab) {
            ps = 0;
            pt += gridDim.x;
            pactive = pt < a.ntiles;
            if (pactive) producer_tile();
        } else {
            pa += (int64_t)RD_BK * a.lda;
#pragma unroll
            for (int u = 0; u < 4; ++u) pb[u] += RD_BK;
        }
    };
    if (pactive) producer_tile();
#pragma unroll 1
    for (int s = 0; s < RD_STAGES - 1 && pactive; ++s) {   // pipeline fill
        issue_half(0);
        issue_half(1);
        finish_slab();
    }
    auto load_frags = [&](int b, int k4, double (&af)[RD_MI], double (&bf)[RD_NJ]) {
        const double *sA = smem + b * RD_STAGE, *sB = sA + RD_A;
#pragma unroll
        for (int i = 0; i < RD_MI; ++i) af[i] = sA[(k4 * 4 + t4) * RD_LDM + wm * (RD_MI * 8) + i * 8 + g];
#pragma unroll
        for (int j = 0; j < RD_NJ; ++j) bf[j] = sB[(wn * (RD_NJ * 8) + j * 8 + g) * RD_LDK + k4 * 4 + t4];
    };

    SmemHist<RD_THREADS, false> H;
    H.h = hist + tid;
    unsigned cg = 0;
    int cur_ty = -1, next_ty = 0;
    double acc[RD_MI][RD_NJ][2];
    double af[2][RD_MI], bf[2][RD_NJ];
    constexpr int NK4 = RD_BK / 4;
    for (int ti = 0; ti < my_tiles; ++ti) {
        const int t = blockIdx.x + ti * gridDim.x;
        const int64_t m0 = (int64_t)(t % a.tiles_m) * RD_BM;
        const int ty = t / a.tiles_m;
#pragma unroll
        for (int i = 0; i < RD_MI; ++i)
#pragma unroll
            for (int j = 0; j < RD_NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        mbar_wait(full + cg % RD_STAGES, (cg / RD_STAGES) & 1u);
        load_frags((int)(cg % RD_STAGES), 0, af[0], bf[0]);
#pragma unroll 1
        for (int s = 0; s < nslab; ++s, ++cg) {
            const int b = (int)(cg % RD_STAGES);
#pragma unroll
            for (int k4 = 0; k4 < NK4; ++k4) {
                if (k4 == NK4 / 2 && pactive) {
                    // refill the buffer released RD_STAGES slabs ago (half a slab of slack since its release)
                    if (pg >= RD_STAGES) mbar_wait(empty + pg % RD_STAGES, (pg / RD_STAGES - 1) & 1u);
                    issue_half(0);
                }
                if (k4 == NK4 / 2 + 1 && pactive) {
                    issue_half(1);
                    finish_slab();
                }
                if (k4 < NK4 - 1) {
                    load_frags(b, k4 + 1, af[(k4 + 1) & 1], bf[(k4 + 1) & 1]);
                } else if (s + 1 < nslab) {
                    mbar_wait(full + (cg + 1) % RD_STAGES, ((cg + 1) / RD_STAGES) & 1u);
                    load_frags((int)((cg + 1) % RD_STAGES), 0, af[0], bf[0]);
                }
#pragma unroll
                for (int i = 0; i < RD_MI; ++i)
#pragma unroll
                    for (int j = 0; j < RD_NJ; ++j)   // D' accumulated: operands swapped (their fragment layouts coincide)
                        dmma884v(acc[i][j][0], acc[i][j][1], bf[k4 & 1][j], af[k4 & 1][i]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + b);   // after the DMMAs that consumed the slab's last fragments
        }

        // ------------------------------------------------------------ epilogue: deposit the tile
        if (ty != cur_ty) {
            if (cur_ty >= 0) {
                publish(cur_ty, true);
                next_ty = cur_ty + 1;
            }
            while (next_ty < ty) publish(next_ty++, false);
            cur_ty = ty;
        }
        const bool sw1 = (t4 & 1) != 0, sw2 = (t4 & 2) != 0;
#pragma unroll
        for (int i = 0; i < RD_MI; ++i) {
            // the 8 rows i*8 .. i*8+7 of this thread's sample
            constexpr int NP = 8;
            double x[NP];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                // w_r = value of sample j = t4 ^ r (two conditional swaps), sent to lane t4 ^ r, which owns that sample
                const double v0 = acc[i][0][e], v1 = acc[i][1][e], v2 = acc[i][2][e], v3 = acc[i][3][e];
                const double u0 = sw1 ? v1 : v0, u1 = sw1 ? v0 : v1, u2 = sw1 ? v3 : v2, u3 = sw1 ? v2 : v3;
                const double w0 = sw2 ? u2 : u0, w1 = sw2 ? u3 : u1, w2 = sw2 ? u0 : u2, w3 = sw2 ? u1 : u3;
                x[4 * e + 0] = w0;
                x[4 * e + 1] = __shfl_xor_sync(0xffffffffu, w1, 1);
                x[4 * e + 2] = __shfl_xor_sync(0xffffffffu, w2, 2);
                x[4 * e + 3] = __shfl_xor_sync(0xffffffffu, w3, 3);
            }
            int c[NP];
            double xi[NP], b[NP][4];
            bool ok = true;   // one branch per group (a branch per particle would fence the particles' instruction streams)
#pragma unroll
            for (int q = 0; q < NP; ++q) ok &= locate_fast(x[q], a.g, c[q], xi[q]);
            if (!ok) {
#pragma unroll
                for (int q = 0; q < NP; ++q) locate(x[q], a.g, c[q], xi[q]);
            }
#pragma unroll
            for (int q = 0; q < NP; ++q) bspline3x6(xi[q], b[q][0], b[q][1], b[q][2], b[q][3]);
            H.add_n<NP>(c, b);
        }
        if (m0 + RD_BM > a.N) {   // ragged last row tile: take the zero rows' 6 B(0) = (1, 4, 1, 0) out again
            const int64_t over = m0 + wm * (RD_MI * 8) + RD_MI * 8 - a.N;   // rows of this warp tile past the matrix
            const double cnt = (double)(over < 0 ? 0 : (over > RD_MI * 8 ? RD_MI * 8 : over));
            H.h[0] -= cnt;
            H.h[RD_THREADS] -= 4.0 * cnt;
            H.h[2 * RD_THREADS] -= cnt;
        }
    }
    cp_async_wait<0>();
    if (cur_ty >= 0) {
        publish(cur_ty, true);
        next_ty = cur_ty + 1;
    }
    const int tn = (a.S + RD_BN - 1) / RD_BN;
    while (next_ty < tn) publish(next_ty++, false);
}

}  // namespace rbm

using namespace rbm;

namespace {

size_t smem_per_sm(const rbm_ctx *ctx)
{
    int per_sm = 0;
    if (cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, ctx->device) != cudaSuccess) return 0;
    return (size_t)per_sm;
}

template <int MI, int STAGES, int CTAS>
int launch_recon(rbm_ctx *ctx, ReconArgs a, int64_t N, int S, int *nchunks)
{
    using Shape = ReconShape<MI, STAGES>;
    const int tm = (int)((N + Shape::BM - 1) / Shape::BM), tn = (S + RD_BN - 1) / RD_BN;
    const int64_t ntiles = (int64_t)tm * tn;
    const size_t smem = Shape::smem(a.g.nh);
    if (CTAS * (smem + 1024) > smem_per_sm(ctx) || ntiles < (int64_t)CTAS * ctx->sm_count || ntiles > (1 << 30)) return RBM_ERR_UNSUPPORTED;
    RBM_CUDA(ctx, cudaFuncSetAttribute(k_recon_deposit<MI, STAGES, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    a.tiles_m = tm; a.ntiles = (int)ntiles;
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)CTAS * ctx->sm_count);
    *nchunks = grid;
    ProfScope ps(ctx, "k_recon_deposit");
    k_recon_deposit<MI, STAGES, CTAS><<<grid, RD_THREADS, smem, ctx->stream>>>(a);
    RBM_LAUNCH_CHECK(ctx);
    return RBM_OK;
}

}  // namespace

// most partials per sample any variant writes (CTAs per SM x SMs)
int rbm_recon_deposit_max_chunks(const rbm_ctx *ctx) { return 2 * ctx->sm_count; }

// part[s][c][bin] (c < *nchunks CTAs) = 6 B-spline deposit of column s of X = A Z, A = N x K at A (lda even, 16-byte
// aligned), Z = K x S (ldz even, 16-byte aligned).  Uniform weights, cubic splines.
// Returns RBM_ERR_UNSUPPORTED (nothing launched) when the shape does not suit the kernel; *nchunks = grid size.
int rbm_recon_deposit(rbm_ctx *ctx, const Geom &g, const double *A, int64_t lda, const double *Z, int64_t ldz, int64_t N,
                      int S, int K, double *part, int *nchunks)
{
    if (g.degree != 3 || K < 1 || (lda & 1) || (ldz & 1) || ((uintptr_t)A & 15) || ((uintptr_t)Z & 15)) return RBM_ERR_UNSUPPORTED;
    ReconArgs a{};
    a.A = A; a.lda = lda; a.Z = Z; a.ldz = ldz; a.N = N; a.S = S; a.K = K; a.g = g; a.part = part;
    // 128 x 64 tiles, two CTAs per SM.  (64 x 64 tiles with three CTAs per SM hide more of the epilogue -- +57 instead of
    // +76 us at N = 1e5, S = 256, K = 173 -- but their two-stage main loop alone is 341 us against 308: measured, not kept.)
    return launch_recon<8, 3, 2>(ctx, a, N, S, nchunks);
}
