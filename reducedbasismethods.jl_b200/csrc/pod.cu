// pod.cu -- snapshot POD (Gram + EVD), basis projection and DEIM behind the C ABI.
// Reference: src/algorithms/evd.jl:26-87, src/eigen.jl:4-7, src/algorithms/deim.jl:6-21.
// The dense contractions run on the DMMA tile kernel of dgemm_kernels.cuh; the small
// symmetric eigenproblem (C x C, C = number of snapshot columns) is cusolverDnDsyevd
// (a plain library call, like LAPACK's `eigen` in the reference).
#include <cusolverDn.h>

#include <algorithm>
#include <cmath>
#include <memory>
#include <numeric>

#include "common.cuh"
#include "dgemm_kernels.cuh"

using namespace rbm;

void rbm_pod_release(rbm_ctx *ctx)
{
    if (ctx->cusolver) {
        cusolverDnDestroy((cusolverDnHandle_t)ctx->cusolver);
        ctx->cusolver = nullptr;
    }
    if (ctx->gemm_work) cudaFree(ctx->gemm_work);
    ctx->gemm_work = nullptr;
    ctx->gemm_work_bytes = 0;
}

namespace {

int get_cusolver(rbm_ctx *ctx, cusolverDnHandle_t *h)
{
    if (!ctx->cusolver) {
        cusolverDnHandle_t hh;
        cusolverStatus_t st = cusolverDnCreate(&hh);
        if (st != CUSOLVER_STATUS_SUCCESS) return rbm_fail(ctx, RBM_ERR_CUDA, "cusolverDnCreate failed (%d)", (int)st);
        ctx->cusolver = hh;
    }
    *h = (cusolverDnHandle_t)ctx->cusolver;
    cusolverDnSetStream(*h, ctx->stream);
    return RBM_OK;
}

// Column blocks -> device-resident blocks + array of column pointers.
struct ColumnSet {
    std::unique_ptr<DevIn[]> staged;  // one per block (host blocks are copied to the device)
    std::vector<const double *> h_cols;
    DevBuf d_cols;
    int64_t ncols = 0;
    bool vec_ok = true;             // every column 16-byte aligned
    const double *uni_base = nullptr;   // all columns are base + c * uni_ld (one block, or blocks adjacent in memory)
    int64_t uni_ld = 0;

    int build(rbm_ctx *ctx, const double *const *blocks, const int64_t *ncols_b, const int64_t *ld, int nblocks,
              int64_t nrows, const char *who)
    {
        if (!blocks || !ncols_b || nblocks < 1 || nrows < 1)
            return rbm_fail(ctx, RBM_ERR_INVALID, "%s: need at least one non-empty column block", who);
        staged.reset(new DevIn[nblocks]);
        for (int b = 0; b < nblocks; ++b) {
            const int64_t ldb = ld ? ld[b] : nrows;
            if (!blocks[b] || ncols_b[b] < 1 || ldb < nrows)
                return rbm_fail(ctx, RBM_ERR_INVALID, "%s: block %d is empty or has ld < nrows (DimensionMismatch)", who, b);
            RBM_TRY(staged[b].stage(ctx, blocks[b], sizeof(double) * (size_t)ldb * (size_t)ncols_b[b]));
            const double *base = staged[b].as<double>();
            for (int64_t c = 0; c < ncols_b[b]; ++c) {
                const double *p = base + (size_t)c * ldb;
                if ((uintptr_t)p & 15) vec_ok = false;
                h_cols.push_back(p);
            }
        }
        ncols = (int64_t)h_cols.size();
        uni_base = h_cols[0];
        uni_ld = ncols > 1 ? (int64_t)(h_cols[1] - h_cols[0]) : nrows;
        for (int64_t c = 1; c < ncols && uni_base; ++c)
            if (h_cols[c] != h_cols[0] + c * uni_ld) uni_base = nullptr;
        RBM_CUDA(ctx, d_cols.alloc(sizeof(double *) * h_cols.size()));
        RBM_CUDA(ctx, cudaMemcpyAsync(d_cols.p, h_cols.data(), sizeof(double *) * h_cols.size(),
                                      cudaMemcpyHostToDevice, ctx->stream));
        return RBM_OK;
    }
    const double *const *dev() const { return d_cols.as<const double *>(); }
};

template <bool TA, bool VEC> int launch_gemm_t(rbm_ctx *ctx, const GemmArgs &a, int ntiles, int nsplit)
{
    constexpr size_t smem = sizeof(double) * GM_STAGES * ((TA ? GM_A_ELEMS_T : GM_A_ELEMS_N) + GM_B_ELEMS) +
                            sizeof(double *) * (GM_BM + GM_BN);   // + the tile's column-pointer table
    RBM_CUDA(ctx, cudaFuncSetAttribute(k_dgemm<TA, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ntiles, nsplit);
    ProfScope ps(ctx, "k_dgemm");
    k_dgemm<TA, VEC><<<grid, GM_THREADS, smem, ctx->stream>>>(a);
    RBM_LAUNCH_CHECK(ctx);
    return RBM_OK;
}

int launch_gemm_mb(rbm_ctx *ctx, const GemmArgs &a, int ntiles, int nsplit)
{
    RBM_CUDA(ctx, cudaFuncSetAttribute(k_dgemm_mb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GM_MB_SMEM));
    dim3 grid(ntiles, nsplit);
    ProfScope ps(ctx, "k_dgemm");
    k_dgemm_mb<<<grid, GM_THREADS, GM_MB_SMEM, ctx->stream>>>(a);
    RBM_LAUNCH_CHECK(ctx);
    return RBM_OK;
}

int launch_gemm_pmb_ab(rbm_ctx *ctx, const GemmArgs &a, int ntiles, int nsplit)
{
    constexpr size_t smem = gm_pmb_smem<false>();
    RBM_CUDA(ctx, cudaFuncSetAttribute(k_dgemm_pmb<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nitems = ntiles * nsplit;
    ProfScope ps(ctx, "k_dgemm");
    k_dgemm_pmb<false><<<std::min(nitems, ctx->sm_count), GM_THREADS, smem, ctx->stream>>>(a, ntiles, nitems);
    RBM_LAUNCH_CHECK(ctx);
    return RBM_OK;
}

// D (M x N, ld ldd, device) = op(A) * B with op(A) = A' (trans_a, contraction over rows K) or A.
// sym: A == B, only the lower-triangular tiles are computed and the result is mirrored.
int run_gemm(rbm_ctx *ctx, bool trans_a, const double *const *acol, const double *const *bcol, int64_t M, int64_t N,
             int64_t K, bool vec, bool sym, double *D, int64_t ldd, const double *a_base = nullptr, int64_t a_ld = 0)
{
    const int tm = (int)((M + GM_BM - 1) / GM_BM), tn = (int)((N + GM_BN - 1) / GM_BN);
    const int ntiles = sym ? tm * (tm + 1) / 2 : tm * tn;   // sym: lower-triangular tiles only (M == N)
    // Split the contraction so that tiles x splits fills whole waves of SMs: pick the split count that
    // minimises (number of waves) x (rows per split), i.e. the idle tail of the last wave.
    int nsplit = 1;
    if (trans_a) {
        const int64_t slabs = (K + GM_BK - 1) / GM_BK;
        const int max_split = (int)std::max<int64_t>(1, std::min<int64_t>(32, slabs / 8));
        double best = 1e300;
        for (int s = 1; s <= max_split; ++s) {
            const int64_t waves = ((int64_t)ntiles * s + ctx->sm_count - 1) / ctx->sm_count;
            const double cost = (double)waves / (double)s * (1.0 + 0.002 * s);   // mild penalty: workspace + reduce
            if (cost < best) { best = cost; nsplit = s; }
        }
    }
    else if (ntiles * 2 <= ctx->sm_count) {
        // small A B products (the reduced operators): a handful of 128 x 128 tiles would run on a handful of SMs,
        // so the contraction is split until the grid covers the machine (at least two slabs per split)
        const int64_t slabs = (K + GM_BK - 1) / GM_BK;
        nsplit = (int)std::max<int64_t>(1, std::min<int64_t>(slabs / 2, ctx->sm_count / ntiles));
    }
    int64_t kps = (K + nsplit - 1) / nsplit;
    kps = ((kps + GM_BK - 1) / GM_BK) * GM_BK;
    nsplit = (int)((K + kps - 1) / kps);
    GemmArgs a{};
    a.acol = acol; a.bcol = bcol; a.M = M; a.N = N; a.K = K; a.k_per_split = kps;
    a.tiles_m = tm; a.same_ab = sym ? 1 : 0;
    a.a_base = trans_a ? nullptr : a_base; a.a_ld = a_ld;
    const bool direct = nsplit == 1 && !sym;
    if (direct) {
        a.out = D; a.ldo = ldd; a.split_stride = 0;
    } else {
        const size_t need = sizeof(double) * (size_t)nsplit * M * N;
        if (ctx->gemm_work_bytes < need) {
            if (ctx->gemm_work) RBM_CUDA(ctx, cudaFree(ctx->gemm_work));
            ctx->gemm_work = nullptr;
            ctx->gemm_work_bytes = 0;
            RBM_CUDA(ctx, cudaMalloc(&ctx->gemm_work, need));
            ctx->gemm_work_bytes = need;
        }
        a.out = (double *)ctx->gemm_work; a.ldo = M; a.split_stride = M * N;
    }
    if (trans_a) {
        if (vec) RBM_TRY(launch_gemm_mb(ctx, a, ntiles, nsplit));   // aligned columns: mbarrier-pipelined kernel
        else RBM_TRY((launch_gemm_t<true, false>(ctx, a, ntiles, nsplit)));
    } else {
        if (vec) RBM_TRY(launch_gemm_pmb_ab(ctx, a, ntiles, nsplit));   // persistent mbarrier-pipelined kernel
        else RBM_TRY((launch_gemm_t<false, false>(ctx, a, ntiles, nsplit)));
    }
    if (!direct) {
        const int64_t n = M * N;
        k_splitk_reduce<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const double *)ctx->gemm_work, nsplit, M * N, M, N, M,
                                                                             D, ldd, sym ? 1 : 0);
        RBM_LAUNCH_CHECK(ctx);
    }
    return RBM_OK;   // asynchronous on ctx->stream (the split-K workspace lives in the context)
}

// upper triangle <- lower triangle (column-major C x C)
__global__ void k_mirror_lower(double *G, int64_t C)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= C * C) return;
    const int64_t i = idx % C, j = idx / C;
    if (i < j) G[j * C + i] = G[i * C + j];
}

// G (device, C x C) = S'S summed over ranks
int gram_device(rbm_ctx *ctx, const ColumnSet &cs, int64_t nrows, double *G)
{
    const int64_t C = cs.ncols;
    RBM_TRY(run_gemm(ctx, true, cs.dev(), cs.dev(), C, C, nrows, cs.vec_ok, true, G, C));
    if (ctx->nccl_comm && ctx->nranks > 1) {
        RBM_TRY(rbm_allreduce_sum_f64(ctx, G, (size_t)C * C));
        // with more than two ranks the reduction order of an element depends on its position in the buffer, so
        // G[i,j] and G[j,i] can differ in the last bit: mirror the lower triangle again (G == G' exactly, on every rank)
        k_mirror_lower<<<(unsigned)((C * C + 255) / 256), 256, 0, ctx->stream>>>(G, C);
        RBM_LAUNCH_CHECK(ctx);
    }
    return RBM_OK;
}

__global__ void k_scale_cols(const double *Om, int64_t C, const int *perm, const double *scale, int k, double *B)
{
    // B[:, j] = Omega[:, perm[j]] * scale[j]  (sign-normalised by the host through scale's sign)
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= C * k) return;
    const int64_t i = idx % C, j = idx / C;
    B[j * C + i] = Om[(size_t)perm[j] * C + i] * scale[j];
}

__global__ void k_col_absmax_sign(const double *Om, int64_t C, const int *perm, int k, double *sign)
{
    // sign[j] = sign of the largest-magnitude entry of Omega[:, perm[j]] (first one on ties)
    const int j = blockIdx.x;
    if (j >= k) return;
    __shared__ double sv[256];
    __shared__ int64_t si[256];
    const double *col = Om + (size_t)perm[j] * C;
    double bv = -1.0;
    int64_t bi = 0;
    for (int64_t i = threadIdx.x; i < C; i += blockDim.x) {
        double v = fabs(col[i]);
        if (v > bv) { bv = v; bi = i; }
    }
    sv[threadIdx.x] = bv; si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            double v2 = sv[threadIdx.x + o]; int64_t i2 = si[threadIdx.x + o];
            if (v2 > sv[threadIdx.x] || (v2 == sv[threadIdx.x] && i2 < si[threadIdx.x])) { sv[threadIdx.x] = v2; si[threadIdx.x] = i2; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) sign[j] = col[si[0]] < 0.0 ? -1.0 : 1.0;
}

}  // namespace

int rbm_make_cols(rbm_ctx *ctx, const double *base, int64_t ld, int64_t ncols, DevBuf &out, bool *vec_ok)
{
    std::vector<const double *> h((size_t)ncols);
    bool ok = true;
    for (int64_t c = 0; c < ncols; ++c) {
        h[c] = base + (size_t)c * ld;
        if ((uintptr_t)h[c] & 15) ok = false;
    }
    if (vec_ok) *vec_ok = ok;
    if (!(out.p && out.bytes >= sizeof(double *) * (size_t)ncols)) {
        RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // a long-lived pointer table may still be in use
        RBM_CUDA(ctx, out.alloc(sizeof(double *) * (size_t)ncols));
    }
    RBM_CUDA(ctx, cudaMemcpyAsync(out.p, h.data(), sizeof(double *) * (size_t)ncols, cudaMemcpyHostToDevice, ctx->stream));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // h is a temporary
    return RBM_OK;
}

int rbm_run_gemm(rbm_ctx *ctx, bool trans_a, const double *const *acol, const double *const *bcol, int64_t M, int64_t N,
                 int64_t K, bool vec, double *D, int64_t ldd, const double *a_base, int64_t a_ld)
{
    return run_gemm(ctx, trans_a, acol, bcol, M, N, K, vec, false, D, ldd, a_base, a_ld);
}

namespace {
__global__ void k_identity(double *A, int n)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n * n) A[idx] = (idx / n == idx % n) ? 1.0 : 0.0;
}
}  // namespace

int rbm_invert(rbm_ctx *ctx, double *A, int n, double *Ainv)
{
    cusolverDnHandle_t h = nullptr;
    RBM_TRY(get_cusolver(ctx, &h));
    int lwork = 0;
    if (cusolverDnDgetrf_bufferSize(h, n, n, A, n, &lwork) != CUSOLVER_STATUS_SUCCESS)
        return rbm_fail(ctx, RBM_ERR_CUDA, "cusolverDnDgetrf_bufferSize failed");
    DevBuf work, piv, info;
    RBM_CUDA(ctx, work.alloc(sizeof(double) * (size_t)std::max(lwork, 1)));
    RBM_CUDA(ctx, piv.alloc(sizeof(int) * (size_t)n));
    RBM_CUDA(ctx, info.alloc(sizeof(int)));
    k_identity<<<(n * n + 255) / 256, 256, 0, ctx->stream>>>(Ainv, n);
    RBM_LAUNCH_CHECK(ctx);
    if (cusolverDnDgetrf(h, n, n, A, n, work.as<double>(), piv.as<int>(), info.as<int>()) != CUSOLVER_STATUS_SUCCESS)
        return rbm_fail(ctx, RBM_ERR_NUMERIC, "cusolverDnDgetrf failed");
    if (cusolverDnDgetrs(h, CUBLAS_OP_N, n, n, A, n, piv.as<int>(), Ainv, n, info.as<int>()) != CUSOLVER_STATUS_SUCCESS)
        return rbm_fail(ctx, RBM_ERR_NUMERIC, "cusolverDnDgetrs failed");
    ctx->launches += 2;
    int hinfo = 0;
    RBM_CUDA(ctx, cudaMemcpyAsync(&hinfo, info.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (hinfo != 0) return rbm_fail(ctx, RBM_ERR_NUMERIC, "singular DEIM interpolation matrix (info = %d) (SingularException)", hinfo);
    return RBM_OK;
}

extern "C" int rbm_gram(rbm_ctx *ctx, const double *const *blocks, const int64_t *ncols, const int64_t *ld, int nblocks,
                        int64_t nrows, double *G)
{
    if (!ctx) return RBM_ERR_INVALID;
    if (!G) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_gram: G == NULL");
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    ColumnSet cs;
    RBM_TRY(cs.build(ctx, blocks, ncols, ld, nblocks, nrows, "rbm_gram"));
    DevOut dG;
    RBM_TRY(dG.prepare(ctx, G, sizeof(double) * (size_t)cs.ncols * cs.ncols));
    RBM_TRY(gram_device(ctx, cs, nrows, dG.as<double>()));
    RBM_TRY(dG.flush(ctx));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RBM_OK;
}

// Decomposition kept between the rank query and the call(s) that form Psi: an explicit, caller-owned handle
// (rbm_pod_factor -> rbm_pod_basis -> rbm_evd_destroy).  Nothing is cached behind the caller's back.
struct rbm_evd {
    rbm_ctx *ctx = nullptr;
    int64_t C = 0, nrows = 0;
    DevBuf eigvec;              // C x C, columns = eigenvectors in the solver's order
    std::vector<double> Ls;     // eigenvalues sorted by |lambda| descending
    std::vector<int> perm;      // sorted position -> solver column
    int k = 0;                  // rank chosen by rbm_pod_factor
};

namespace {

// Gram + eigen + sorteigen + rank (evd.jl:32-46 / :65-78)
int evd_factor(rbm_ctx *ctx, const ColumnSet &cs, int64_t nrows, int k, double tolerance, rbm_evd &f)
{
    const int64_t C = cs.ncols;
    if (k > C) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pod_evd: k = %d exceeds the %lld snapshot columns", k, (long long)C);
    if (C > 32768) return rbm_fail(ctx, RBM_ERR_UNSUPPORTED, "rbm_pod_evd: %lld columns (limit 32768)", (long long)C);
    f.ctx = ctx; f.C = C; f.nrows = nrows;
    // ---- Gram matrix (evd.jl:33 / :65) and eigen(G) (evd.jl:35 / :67): symmetric path, eigenvectors overwrite G
    DevBuf &G = f.eigvec;
    std::vector<double> lam(C);
    RBM_CUDA(ctx, G.alloc(sizeof(double) * (size_t)C * C));
    RBM_TRY(gram_device(ctx, cs, nrows, G.as<double>()));
    cusolverDnHandle_t h = nullptr;
    RBM_TRY(get_cusolver(ctx, &h));
    DevBuf dLam, dInfo, dWork;
    RBM_CUDA(ctx, dLam.alloc(sizeof(double) * C));
    RBM_CUDA(ctx, dInfo.alloc(sizeof(int)));
    int lwork = 0;
    cusolverStatus_t st = cusolverDnDsyevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)C,
                                                      G.as<double>(), (int)C, dLam.as<double>(), &lwork);
    if (st != CUSOLVER_STATUS_SUCCESS) return rbm_fail(ctx, RBM_ERR_CUDA, "cusolverDnDsyevd_bufferSize failed (%d)", (int)st);
    RBM_CUDA(ctx, dWork.alloc(sizeof(double) * (size_t)lwork));
    st = cusolverDnDsyevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)C, G.as<double>(), (int)C,
                          dLam.as<double>(), dWork.as<double>(), lwork, dInfo.as<int>());
    ctx->launches++;
    if (st != CUSOLVER_STATUS_SUCCESS) return rbm_fail(ctx, RBM_ERR_NUMERIC, "cusolverDnDsyevd failed (%d)", (int)st);
    int info = 0;
    RBM_CUDA(ctx, cudaMemcpyAsync(lam.data(), dLam.p, sizeof(double) * C, cudaMemcpyDeviceToHost, ctx->stream));
    RBM_CUDA(ctx, cudaMemcpyAsync(&info, dInfo.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (info != 0) return rbm_fail(ctx, RBM_ERR_NUMERIC, "rbm_pod_evd: eigensolver did not converge (info = %d)", info);

    // ---- sorteigen: by |lambda| descending, stable (src/eigen.jl:4-7)
    f.perm.resize(C);
    std::iota(f.perm.begin(), f.perm.end(), 0);
    std::stable_sort(f.perm.begin(), f.perm.end(), [&](int a, int b) { return std::fabs(lam[a]) > std::fabs(lam[b]); });
    f.Ls.resize(C);
    for (int64_t i = 0; i < C; ++i) f.Ls[i] = lam[f.perm[i]];
    // ---- rank from the energy criterion (evd.jl:38-46 / :70-78)
    double E = 0.0;
    for (int64_t i = 0; i < C; ++i) E += f.Ls[i];
    // An all-zero (or non-finite) snapshot block has no POD basis: the reference would divide 0/0 in the rank loop and
    // scale by 1/sqrt(0); report it instead of handing NaN columns to DEIM.
    if (!(E > 0.0) || !std::isfinite(E))
        return rbm_fail(ctx, RBM_ERR_NUMERIC, "rbm_pod_evd: snapshot energy sum(Lambda) = %g is not positive and finite", E);
    int kk = k;
    if (kk == 0) {
        double Er = 0.0;
        kk = 1;
        while (kk < C && 1.0 - tolerance > (Er + f.Ls[kk - 1]) / E) {
            Er += f.Ls[kk - 1];
            ++kk;
        }
    }
    for (int j = 0; j < kk; ++j)
        if (!(std::fabs(f.Ls[j]) > 0.0))
            return rbm_fail(ctx, RBM_ERR_NUMERIC, "rbm_pod_evd: eigenvalue %d of the %d requested modes is zero (rank-deficient snapshots)", j + 1, kk);
    f.k = kk;
    return RBM_OK;
}

// Psi = S * Omega[:,1:k] * diag(|Lambda|^-1/2)  (evd.jl:49 / :81).  The sign of each eigenvector is
// solver-dependent; it is fixed here by making its largest-magnitude entry positive.
int evd_basis(rbm_ctx *ctx, const rbm_evd &f, const ColumnSet &cs, int64_t nrows, int kk, double *Psi, int64_t ldpsi)
{
    const int64_t C = f.C;
    DevBuf dPerm, dScale, dB, dBcols;
    RBM_CUDA(ctx, dPerm.alloc(sizeof(int) * kk));
    RBM_CUDA(ctx, dScale.alloc(sizeof(double) * kk));
    RBM_CUDA(ctx, dB.alloc(sizeof(double) * (size_t)C * kk));
    RBM_CUDA(ctx, cudaMemcpyAsync(dPerm.p, f.perm.data(), sizeof(int) * kk, cudaMemcpyHostToDevice, ctx->stream));
    k_col_absmax_sign<<<kk, 256, 0, ctx->stream>>>(f.eigvec.as<double>(), C, dPerm.as<int>(), kk, dScale.as<double>());
    RBM_LAUNCH_CHECK(ctx);
    std::vector<double> sc(kk);
    RBM_CUDA(ctx, cudaMemcpyAsync(sc.data(), dScale.p, sizeof(double) * kk, cudaMemcpyDeviceToHost, ctx->stream));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int j = 0; j < kk; ++j) sc[j] *= 1.0 / std::sqrt(std::fabs(f.Ls[j]));
    RBM_CUDA(ctx, cudaMemcpyAsync(dScale.p, sc.data(), sizeof(double) * kk, cudaMemcpyHostToDevice, ctx->stream));
    {
        const int64_t n = C * kk;
        k_scale_cols<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(f.eigvec.as<double>(), C, dPerm.as<int>(),
                                                                          dScale.as<double>(), kk, dB.as<double>());
        RBM_LAUNCH_CHECK(ctx);
    }
    std::vector<const double *> bcols(kk);
    for (int j = 0; j < kk; ++j) bcols[j] = dB.as<double>() + (size_t)j * C;
    RBM_CUDA(ctx, dBcols.alloc(sizeof(double *) * kk));
    RBM_CUDA(ctx, cudaMemcpyAsync(dBcols.p, bcols.data(), sizeof(double *) * kk, cudaMemcpyHostToDevice, ctx->stream));
    const bool psi_dev = rbm_is_device_ptr(Psi);
    int64_t ldp = psi_dev ? ldpsi : nrows;
    DevBuf psi_stage;
    double *dpsi = Psi;
    if (!psi_dev) {
        RBM_CUDA(ctx, psi_stage.alloc(sizeof(double) * (size_t)nrows * kk));
        dpsi = psi_stage.as<double>();
    }
    const bool vec = cs.vec_ok && (C % 2 == 0);
    RBM_TRY(run_gemm(ctx, false, cs.dev(), dBcols.as<const double *>(), nrows, kk, C, vec, false, dpsi, ldp, cs.uni_base, cs.uni_ld));
    if (!psi_dev)
        RBM_CUDA(ctx, cudaMemcpy2DAsync(Psi, sizeof(double) * ldpsi, dpsi, sizeof(double) * nrows, sizeof(double) * nrows,
                                        kk, cudaMemcpyDeviceToHost, ctx->stream));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RBM_OK;
}

}  // namespace

extern "C" int rbm_pod_factor(rbm_ctx *ctx, const double *const *blocks, const int64_t *ncols, const int64_t *ld,
                              int nblocks, int64_t nrows, int k, double tolerance, double *Lambda, int *k_out,
                              rbm_evd **out)
{
    if (!ctx) return RBM_ERR_INVALID;
    if (!k_out || !out) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pod_factor: k_out / out == NULL");
    *out = nullptr;
    if (k < 0) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pod_factor: k < 0");
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    ColumnSet cs;
    RBM_TRY(cs.build(ctx, blocks, ncols, ld, nblocks, nrows, "rbm_pod_factor"));
    std::unique_ptr<rbm_evd> f(new (std::nothrow) rbm_evd());
    if (!f) return rbm_fail(ctx, RBM_ERR_NOMEM, "rbm_pod_factor: out of host memory");
    RBM_TRY(evd_factor(ctx, cs, nrows, k, tolerance, *f));
    *k_out = f->k;
    if (Lambda) RBM_CUDA(ctx, cudaMemcpyAsync(Lambda, f->Ls.data(), sizeof(double) * f->C, cudaMemcpyDefault, ctx->stream));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = f.release();
    return RBM_OK;
}

extern "C" int rbm_pod_basis(rbm_evd *f, const double *const *blocks, const int64_t *ncols, const int64_t *ld, int nblocks,
                             int64_t nrows, int k, double *Psi, int64_t ldpsi)
{
    if (!f) return RBM_ERR_INVALID;
    rbm_ctx *ctx = f->ctx;
    if (!Psi || ldpsi < nrows) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pod_basis: Psi == NULL or ldpsi < nrows");
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    ColumnSet cs;
    RBM_TRY(cs.build(ctx, blocks, ncols, ld, nblocks, nrows, "rbm_pod_basis"));
    // the blocks may be the same matrix (the POD basis) or any matrix with the same columns (e.g. this rank's rows)
    if (cs.ncols != f->C)
        return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pod_basis: %lld columns, the decomposition has %lld (DimensionMismatch)",
                        (long long)cs.ncols, (long long)f->C);
    const int kk = k > 0 ? k : f->k;
    if (kk > f->C) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pod_basis: k = %d exceeds the %lld columns", kk, (long long)f->C);
    for (int j = 0; j < kk; ++j)
        if (!(std::fabs(f->Ls[j]) > 0.0)) return rbm_fail(ctx, RBM_ERR_NUMERIC, "rbm_pod_basis: eigenvalue %d is zero", j + 1);
    return evd_basis(ctx, *f, cs, nrows, kk, Psi, ldpsi);
}

extern "C" void rbm_evd_destroy(rbm_evd *f)
{
    if (!f) return;
    cudaSetDevice(f->ctx->device);
    cudaStreamSynchronize(f->ctx->stream);
    delete f;
}

extern "C" int rbm_pod_evd(rbm_ctx *ctx, const double *const *blocks, const int64_t *ncols, const int64_t *ld,
                           int nblocks, int64_t nrows, int k, double tolerance, double *Lambda, int *k_out,
                           double *Psi, int64_t ldpsi, int psi_capacity)
{
    if (!ctx) return RBM_ERR_INVALID;
    if (!k_out) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pod_evd: k_out == NULL");
    if (k < 0) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pod_evd: k < 0");
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    ColumnSet cs;
    RBM_TRY(cs.build(ctx, blocks, ncols, ld, nblocks, nrows, "rbm_pod_evd"));
    if (Psi && ldpsi < nrows) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pod_evd: ldpsi < nrows");
    rbm_evd f;   // lives for this call only: a rank query (Psi == NULL) keeps nothing (use rbm_pod_factor for that)
    RBM_TRY(evd_factor(ctx, cs, nrows, k, tolerance, f));
    *k_out = f.k;
    if (Lambda) RBM_CUDA(ctx, cudaMemcpyAsync(Lambda, f.Ls.data(), sizeof(double) * f.C, cudaMemcpyDefault, ctx->stream));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (!Psi) return RBM_OK;
    if (k == 0 && f.k > psi_capacity)
        return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pod_evd: rank %d exceeds psi_capacity %d (use rbm_pod_factor / rbm_pod_basis)", f.k, psi_capacity);
    return evd_basis(ctx, f, cs, nrows, f.k, Psi, ldpsi);
}

extern "C" int rbm_project(rbm_ctx *ctx, const double *Psi, int64_t n, int64_t ldpsi, int k, const double *S, int64_t lds,
                           int64_t m, double *Z, int64_t ldz)
{
    if (!ctx) return RBM_ERR_INVALID;
    if (!Psi || !S || !Z || n < 1 || k < 1 || m < 1 || ldpsi < n || lds < n || ldz < k)
        return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_project: bad argument (DimensionMismatch)");
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    ColumnSet a, b;
    const int64_t kc = k, mc = m;
    RBM_TRY(a.build(ctx, &Psi, &kc, &ldpsi, 1, n, "rbm_project"));
    RBM_TRY(b.build(ctx, &S, &mc, &lds, 1, n, "rbm_project"));
    DevBuf dZ;
    RBM_CUDA(ctx, dZ.alloc(sizeof(double) * (size_t)k * m));
    RBM_TRY(run_gemm(ctx, true, a.dev(), b.dev(), k, m, n, a.vec_ok && b.vec_ok, false, dZ.as<double>(), k));
    RBM_TRY(rbm_allreduce_sum_f64(ctx, dZ.as<double>(), (size_t)k * m));
    RBM_CUDA(ctx, cudaMemcpy2DAsync(Z, sizeof(double) * ldz, dZ.p, sizeof(double) * k, sizeof(double) * k, m,
                                    cudaMemcpyDefault, ctx->stream));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RBM_OK;
}

// ------------------------------------------------------------------------ DEIM
//
// deim.jl:11-19 recomputes  c = (Pi'Psi_{1:i-1}) \ (Pi'Psi_i),  r = Psi_i - Psi_{1:i-1} c  from
// scratch every iteration (O(N k^3) overall).  r is the unique element of
// Psi_i + span(Psi_1..Psi_{i-1}) that vanishes at the selected rows, so it can equally be
// written in the basis of the previous residuals r_1..r_{i-1} (which span the same space
// and are "lower triangular" at the selected rows): r_i = Psi_i - sum_j d_j r_j with d from a
// forward substitution.  The residuals overwrite a working copy of Psi column by column.

namespace {

// d = forward substitution:  T d = b,  T[a][j] = W[idx_a, j] (j <= a), b[a] = W[idx_a, i].
// T is stored column-major (Tt[j * ldt + a] = T[a][j]) together with the reciprocals of its diagonal.
// Single block, blocked by 32 columns: warp 0 solves the 32 x 32 diagonal block from shared memory with shuffles (no
// block barrier and no global load per column), then every thread updates whole rows below the block with the block's
// 32 coefficients in ascending column order (coalesced reads of Tt).
// Wcol != NULL (single rank): b is gathered here from column i of W at the selected rows instead of a separate kernel.
__global__ void __launch_bounds__(256) k_deim_coeffs(const double *Tt, const double *rdiag, int ldt, int i, double *d,
                                                    const double *bvec, const double *Wcol, const int64_t *idx)
{
    extern __shared__ double sh[];
    double *b = sh;          // [i] right-hand side, updated in place
    __shared__ double dblk[32];
    __shared__ double tb[32][33];   // the diagonal block, tb[a][j] = T[j0+a][j0+j]
    const int tid = threadIdx.x, lane = tid & 31;
    for (int a = tid; a < i; a += blockDim.x) b[a] = Wcol ? Wcol[idx[a]] : bvec[a];
    __syncthreads();
    for (int j0 = 0; j0 < i; j0 += 32) {
        const int nb = min(32, i - j0);
        for (int q = tid; q < nb * nb; q += blockDim.x) {
            const int j = q / nb, a = q - j * nb;
            tb[a][j] = j <= a ? Tt[(size_t)(j0 + j) * ldt + j0 + a] : 0.0;
        }
        __syncthreads();
        if (tid < 32) {
            double ba = lane < nb ? b[j0 + lane] : 0.0;
            const double rd = lane < nb ? rdiag[j0 + lane] : 0.0;
            for (int j = 0; j < nb; ++j) {
                double dj = 0.0;
                if (lane == j) dj = ba * rd;
                dj = __shfl_sync(0xffffffffu, dj, j);
                if (lane == j) dblk[j] = dj;
                if (lane > j && lane < nb) ba -= dj * tb[lane][j];
            }
        }
        __syncthreads();
        for (int j = tid; j < nb; j += blockDim.x) d[j0 + j] = dblk[j];
        if (nb == 32) {
            // all 32 loads of a row are independent of the running sum: issue them together (one L2 round trip per
            // row instead of one per unrolled group), then accumulate in ascending column order
            for (int a = j0 + 32 + tid; a < i; a += blockDim.x) {
                const double *tcol = Tt + (size_t)j0 * ldt + a;
                double tv[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) tv[j] = tcol[(size_t)j * ldt];
                double ba = b[a];
#pragma unroll
                for (int j = 0; j < 32; ++j) ba -= dblk[j] * tv[j];
                b[a] = ba;
            }
        }   // nb < 32: last block, no rows below it
        __syncthreads();
    }
}

// gather W[row, col] for the (locally owned) selected rows; rows owned elsewhere contribute 0
__global__ void k_deim_gather_rows(const double *W, int64_t ld, const int64_t *idx_global, int count, int64_t row0,
                                   int64_t nloc, int col, double *out)
{
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= count) return;
    int64_t r = idx_global[a] - row0;
    out[a] = (r >= 0 && r < nloc) ? W[(size_t)col * ld + r] : 0.0;
}

// gather row `row` of W, columns 0..i (for the triangular table); zero when not owned
__global__ void k_deim_gather_cols(const double *W, int64_t ld, const double *pair, int64_t row0, int64_t nloc, int ncols,
                                   double *out)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    int64_t r = (int64_t)pair[1] - row0;
    out[j] = (r >= 0 && r < nloc) ? W[(size_t)j * ld + r] : 0.0;
}

// r = W[:,i] - W[:,0:i] d  (in place), block-level signed arg-max (first index on ties).
// A block owns DR_ROWS = 64 rows at a time (two per lane, one 16-byte load per column and lane: 512 contiguous bytes per
// warp) and its 8 warps split the columns j = warp, warp + 8, ...; the eight partial sums of a row are combined in warp
// order.  Small N (the reference's sizes: W lives in L2) is spread over all SMs this way, and at large N every warp
// keeps four independent 512-byte loads in flight (HBM-bound: 8 N (i + 2) bytes per call).
constexpr int DR_ROWS = 64;
template <bool SPLIT>
__global__ void __launch_bounds__(256) k_deim_residual(double *W, int64_t ld, int64_t n, int i, const double *d,
                                                      double *blk_val, int64_t *blk_idx, int vec)
{
    extern __shared__ double sd[];  // d[i]
    __shared__ double part[SPLIT ? 8 : 1][DR_ROWS];
    __shared__ double sv[256];
    __shared__ int64_t si[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j = tid; j < i; j += blockDim.x) sd[j] = d[j];
    __syncthreads();
    double bv = -INFINITY;
    int64_t bi = INT64_MAX;
    const int64_t nrb = (n + DR_ROWS - 1) / DR_ROWS;
    // SPLIT: the block's 8 warps share one 64-row slab and split the columns; else every warp owns its own slab
    const int64_t rb0 = SPLIT ? (int64_t)blockIdx.x : (int64_t)blockIdx.x * 8 + warp;
    const int64_t rbs = SPLIT ? (int64_t)gridDim.x : (int64_t)gridDim.x * 8;
    const int jfirst = SPLIT ? warp : 0, jstep = SPLIT ? 8 : 1;
    for (int64_t rb = rb0; rb < nrb; rb += rbs) {
        const int64_t r = rb * DR_ROWS + 2 * lane;
        const bool ok0 = r < n, ok1 = r + 1 < n;
        double a0 = 0.0, a1 = 0.0;
        if (vec && ok1) {
            int j = jfirst;
            for (; j + 3 * jstep < i; j += 4 * jstep) {   // four independent 512-byte loads per warp in flight
                const double2 w0 = *reinterpret_cast<const double2 *>(W + (size_t)j * ld + r);
                const double2 w1 = *reinterpret_cast<const double2 *>(W + (size_t)(j + jstep) * ld + r);
                const double2 w2 = *reinterpret_cast<const double2 *>(W + (size_t)(j + 2 * jstep) * ld + r);
                const double2 w3 = *reinterpret_cast<const double2 *>(W + (size_t)(j + 3 * jstep) * ld + r);
                a0 = fma(sd[j], w0.x, a0); a1 = fma(sd[j], w0.y, a1);
                a0 = fma(sd[j + jstep], w1.x, a0); a1 = fma(sd[j + jstep], w1.y, a1);
                a0 = fma(sd[j + 2 * jstep], w2.x, a0); a1 = fma(sd[j + 2 * jstep], w2.y, a1);
                a0 = fma(sd[j + 3 * jstep], w3.x, a0); a1 = fma(sd[j + 3 * jstep], w3.y, a1);
            }
            for (; j < i; j += jstep) {
                const double2 w0 = *reinterpret_cast<const double2 *>(W + (size_t)j * ld + r);
                a0 = fma(sd[j], w0.x, a0); a1 = fma(sd[j], w0.y, a1);
            }
        } else {
            for (int j = jfirst; j < i; j += jstep) {
                if (ok0) a0 = fma(sd[j], W[(size_t)j * ld + r], a0);
                if (ok1) a1 = fma(sd[j], W[(size_t)j * ld + r + 1], a1);
            }
        }
        if (SPLIT) {
            part[warp][2 * lane] = a0;
            part[warp][2 * lane + 1] = a1;
            __syncthreads();
            if (tid < DR_ROWS) {
                const int64_t rr = rb * DR_ROWS + tid;
                if (rr < n) {
                    double s = 0.0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) s += part[q][tid];
                    const double acc = W[(size_t)i * ld + rr] - s;
                    if (i > 0) W[(size_t)i * ld + rr] = acc;
                    if (acc > bv) { bv = acc; bi = rr; }
                }
            }
            __syncthreads();
        } else {
            if (ok0) {
                const double acc = W[(size_t)i * ld + r] - a0;
                if (i > 0) W[(size_t)i * ld + r] = acc;
                if (acc > bv) { bv = acc; bi = r; }
            }
            if (ok1) {
                const double acc = W[(size_t)i * ld + r + 1] - a1;
                if (i > 0) W[(size_t)i * ld + r + 1] = acc;
                if (acc > bv) { bv = acc; bi = r + 1; }
            }
        }
    }
    sv[tid] = bv; si[tid] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) {
            const double v2 = sv[tid + o]; const int64_t i2 = si[tid + o];
            if (v2 > sv[tid] || (v2 == sv[tid] && i2 < si[tid])) { sv[tid] = v2; si[tid] = i2; }
        }
        __syncthreads();
    }
    if (tid == 0) { blk_val[blockIdx.x] = sv[0]; blk_idx[blockIdx.x] = si[0]; }
}

// block-wide signed arg-max over the per-block candidates (first index on ties); result in sv[0], si[0]
__device__ __forceinline__ void deim_pick_block(const double *blk_val, const int64_t *blk_idx, int nblk, double *sv, int64_t *si)
{
    const int tid = threadIdx.x;
    double bv = -INFINITY;
    int64_t bi = INT64_MAX;
    for (int b = tid; b < nblk; b += blockDim.x)
        if (blk_val[b] > bv || (blk_val[b] == bv && blk_idx[b] < bi)) { bv = blk_val[b]; bi = blk_idx[b]; }
    sv[tid] = bv; si[tid] = bi;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (tid < o) {
            const double v2 = sv[tid + o]; const int64_t i2 = si[tid + o];
            if (v2 > sv[tid] || (v2 == sv[tid] && i2 < si[tid])) { sv[tid] = v2; si[tid] = i2; }
        }
        __syncthreads();
    }
    // a column without any comparable entry (all NaN): every `>` above was false.  Select row 0 so that the row read
    // that follows stays in bounds; rbm_deim then reports the non-finite pivot as RBM_ERR_NUMERIC.
    if (tid == 0 && si[0] == INT64_MAX) si[0] = 0;
    __syncthreads();
}

__global__ void __launch_bounds__(256) k_deim_pick(const double *blk_val, const int64_t *blk_idx, int nblk, int64_t row0, double *pair)
{
    __shared__ double sv[256];
    __shared__ int64_t si[256];
    deim_pick_block(blk_val, blk_idx, nblk, sv, si);
    if (threadIdx.x == 0) {
        pair[0] = sv[0];
        pair[1] = (double)(si[0] + row0);  // global row (exact for rows < 2^53)
    }
}

// row i of the triangular table (stored column-major) and the reciprocal of its diagonal entry
__global__ void k_deim_commit(const double *pair, const double *rowvals, int i, int64_t *idx_global, double *Tt, double *rdiag,
                              int ldt)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) idx_global[i] = (int64_t)pair[1];
    if (j <= i) Tt[(size_t)j * ldt + i] = rowvals[j];
    if (j == i) rdiag[i] = 1.0 / rowvals[i];
}

// single rank: pick the new point, read row `point` of the residual columns 0..i and commit it, in one block
__global__ void __launch_bounds__(256) k_deim_pick_commit(const double *blk_val, const int64_t *blk_idx, int nblk, const double *W,
                                                         int64_t ld, int i, int64_t *idx_global, double *Tt, double *rdiag,
                                                         int ldt)
{
    __shared__ double sv[256];
    __shared__ int64_t si[256];
    deim_pick_block(blk_val, blk_idx, nblk, sv, si);
    const int64_t r = si[0];
    if (threadIdx.x == 0) idx_global[i] = r;
    for (int j = threadIdx.x; j <= i; j += blockDim.x) {
        const double v = W[(size_t)j * ld + r];
        Tt[(size_t)j * ldt + i] = v;
        if (j == i) rdiag[i] = 1.0 / v;
    }
}


// ------------------------------------------------------------------ single-rank DEIM: blocked, persistent, cooperative
// The selection is a sequence of k dependent steps (new residual -> arg-max -> new point).  Launching three kernels per
// point made the reference's own size (N = 1e4, k = 724) pure launch latency (63 ms), and the left-looking residual
// re-read all previous columns for every point (4 N k^2 bytes).  Here ONE cooperative kernel (one CTA per SM, rows split
// over the CTAs) walks the points with one grid barrier each, blocked like a right-looking LU:
//   * inside a panel of NB columns the residual of column i only needs the panel's previous residuals
//     (coefficients from a <= 31 x 31 forward substitution done by one warp with shuffles, redundantly in every CTA);
//   * after a panel, the trailing columns are updated once: W[:, j] -= R_panel * (T_panel^-1 W[P_panel, j]), so that they
//     vanish at the panel's points (coefficients computed by the warps of all CTAs, one trailing column per warp).
// Same points as deim.jl:6-21 (signed arg-max, first index on ties).  Bytes: ~ 8 N (NB k / 2 + k^2 / NB) instead of 4 N k^2.
// Between two grid barriers a CTA only writes rows it owns, of buffers nobody gathers from until the next barrier.
struct DeimArgs {
    double *W;            // n x k working copy (ld = n): columns receive the trailing updates of earlier panels
    double *R;            // n x 32: residuals of the current panel.  Kept apart from W so that the columns of W are
                          // read-only while a panel is processed: a CTA that runs ahead and stores its rows of residual
                          // i must not change what a slower CTA still gathers from column i at the panel's points.
    int64_t n;
    int k;
    int64_t rows_per_cta; // even
    double *cand_val;     // [2][gridDim] per-CTA arg-max candidates, double-buffered by point parity
    int64_t *cand_idx;
    double *D;            // [k][32] trailing-update coefficients of the current panel
    int64_t *idx;         // [k] selected rows
    double *piv;          // [k] pivots r_i(p_i) (checked on the host)
    unsigned *bar;        // {arrivals, generation}
};

__device__ __forceinline__ void deim_grid_sync(unsigned *bar, unsigned nblocks)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned *gen = bar + 1;
        const unsigned g = *gen;
        __threadfence();
        if (atomicAdd(bar, 1u) == nblocks - 1) {
            bar[0] = 0;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (*gen == g) { }
        }
        __threadfence();
    }
    __syncthreads();
}

// block-wide signed arg-max of (bv, bi) (first index on ties); result broadcast in sv[0], si[0]
template <int THREADS>
__device__ __forceinline__ void deim_block_argmax(double bv, int64_t bi, double *sv, int64_t *si)
{
    const int tid = threadIdx.x;
    sv[tid] = bv; si[tid] = bi;
    __syncthreads();
    for (int o = THREADS / 2; o > 0; o >>= 1) {
        if (tid < o) {
            const double v2 = sv[tid + o]; const int64_t i2 = si[tid + o];
            if (v2 > sv[tid] || (v2 == sv[tid] && i2 < si[tid])) { sv[tid] = v2; si[tid] = i2; }
        }
        __syncthreads();
    }
}

// THREADS = 256 with NB = 32 (matrix in L2: latency-bound, few warps suffice); 512 with NB = 16 (matrix streams from HBM:
// 512 threads x 8 independent 16-byte loads keep ~64 KB per SM in flight)
template <int NB, bool VEC, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_deim_blocked(const DeimArgs a)
{
    __shared__ double tpp[32][33];     // tpp[p][j] = residual j of the panel at the panel's point p (j <= p)
    __shared__ double rd[32], dco[32];
    __shared__ int64_t pts[32];
    __shared__ double sv[THREADS];
    __shared__ int64_t si[THREADS];
    __shared__ double Dt[64][NB];      // coefficient tile of the trailing update
    constexpr int NW = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned nblk = gridDim.x;
    const int64_t n = a.n, ld = a.n;
    const int k = a.k;
    double *W = a.W;
    double *R = a.R;
    const int64_t r0 = min(n, (int64_t)blockIdx.x * a.rows_per_cta), r1 = min(n, r0 + a.rows_per_cta);
    constexpr int RSTEP = VEC ? 2 * THREADS : THREADS;     // rows per block sweep (VEC: two rows per thread)
    for (int j0 = 0; j0 < k; j0 += NB) {
        const int nbp = min(NB, k - j0);
        for (int l = 0; l < nbp; ++l) {
            const int i = j0 + l;
            double *Wi = W + (size_t)i * ld;
            // ---- coefficients of column i in the panel's previous residuals (forward substitution, warp 0)
            if (l > 0 && warp == 0) {
                double b = lane < l ? __ldcg(Wi + pts[lane]) : 0.0;
                for (int j = 0; j < l; ++j) {
                    const double dj = __shfl_sync(0xffffffffu, b, j) * rd[j];
                    if (lane == j) dco[j] = dj;
                    if (lane > j && lane < l) b -= dj * tpp[lane][j];
                }
            }
            __syncthreads();
            // ---- residual of column i on this CTA's rows, local signed arg-max
            double bv = -INFINITY;
            int64_t bi = INT64_MAX;
            const double *Wp = R;              // previous residuals of the panel
            double *Ri = R + (size_t)l * ld;   // this point's residual
            if (VEC) {
                // one row pair: cur = W_i - sum_j d_j R_j, eight independent 16-byte loads in flight
                auto residual_pair = [&](int64_t r) -> double2 {
                    double a0 = 0.0, a1 = 0.0;
                    const double2 cur0 = *reinterpret_cast<const double2 *>(Wi + r);
                    int j = 0;
                    for (; j + 7 < l; j += 8) {
                        double2 w[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) w[u] = *reinterpret_cast<const double2 *>(Wp + (size_t)(j + u) * ld + r);
#pragma unroll
                        for (int u = 0; u < 8; ++u) { a0 = fma(dco[j + u], w[u].x, a0); a1 = fma(dco[j + u], w[u].y, a1); }
                    }
                    {   // the last up-to-seven columns, requested together as well
                        double2 w[7];
#pragma unroll
                        for (int u = 0; u < 7; ++u) w[u] = j + u < l ? *reinterpret_cast<const double2 *>(Wp + (size_t)(j + u) * ld + r) : make_double2(0.0, 0.0);
#pragma unroll
                        for (int u = 0; u < 7; ++u) if (j + u < l) { a0 = fma(dco[j + u], w[u].x, a0); a1 = fma(dco[j + u], w[u].y, a1); }
                    }
                    return make_double2(cur0.x - a0, cur0.y - a1);
                };
                auto commit_pair = [&](int64_t r, const double2 cur) {
                    *reinterpret_cast<double2 *>(Ri + r) = cur;
                    if (cur.x > bv) { bv = cur.x; bi = r; }
                    if (cur.y > bv) { bv = cur.y; bi = r + 1; }
                };
                int64_t r = r0 + 2 * tid;
                if (l < 8) {
                    // few columns: one sweep has at most eight loads per thread in flight, which does not cover the HBM latency;
                    // two independent sweeps at a time (ascending r within a thread: first index on ties is unchanged)
                    for (; r + RSTEP < r1; r += 2 * RSTEP) {
                        const double2 c0 = residual_pair(r), c1 = residual_pair(r + RSTEP);
                        commit_pair(r, c0);
                        commit_pair(r + RSTEP, c1);
                    }
                }
                for (; r < r1; r += RSTEP) commit_pair(r, residual_pair(r));
            } else {
                for (int64_t r = r0 + tid; r < r1; r += RSTEP) {
                    double a0 = 0.0;
                    for (int j = 0; j < l; ++j) a0 = fma(dco[j], Wp[(size_t)j * ld + r], a0);
                    const double cur = Wi[r] - a0;
                    Ri[r] = cur;
                    if (cur > bv) { bv = cur; bi = r; }
                }
            }
            deim_block_argmax<THREADS>(bv, bi, sv, si);
            if (tid == 0) {
                a.cand_val[(size_t)(i & 1) * nblk + blockIdx.x] = sv[0];
                a.cand_idx[(size_t)(i & 1) * nblk + blockIdx.x] = si[0];
            }
            deim_grid_sync(a.bar, nblk);
            // ---- the new point: arg-max over the CTAs' candidates (every CTA, redundantly)
            bv = -INFINITY; bi = INT64_MAX;
            for (unsigned c = tid; c < nblk; c += THREADS) {
                const double v2 = __ldcg(a.cand_val + (size_t)(i & 1) * nblk + c);
                const int64_t i2 = __ldcg(a.cand_idx + (size_t)(i & 1) * nblk + c);
                if (v2 > bv || (v2 == bv && i2 < bi)) { bv = v2; bi = i2; }
            }
            deim_block_argmax<THREADS>(bv, bi, sv, si);
            // a column without any comparable entry (all NaN): row 0 keeps the reads in bounds; the host reports the pivot
            const int64_t p = si[0] == INT64_MAX ? 0 : si[0];
            __syncthreads();
            if (tid == 0) {
                pts[l] = p;
                if (blockIdx.x == 0) a.idx[i] = p;
            }
            if (warp == 0 && lane <= l) {
                const double v = __ldcg(R + (size_t)lane * ld + p);   // residuals j0..i at the new point
                tpp[l][lane] = v;
                if (lane == l) {
                    rd[l] = 1.0 / v;
                    if (blockIdx.x == 0) a.piv[i] = v;
                }
            }
            __syncthreads();
        }
        // ---- trailing update: W[:, j] -= R_panel * (T_panel^-1 W[P_panel, j]) for j beyond the panel
        const int ntrail = k - (j0 + nbp);
        if (ntrail <= 0) break;
        for (int jt = blockIdx.x * NW + warp; jt < ntrail; jt += nblk * NW) {   // one trailing column per warp
            const double *Wj = W + (size_t)(j0 + nbp + jt) * ld;
            double b = lane < nbp ? __ldcg(Wj + pts[lane]) : 0.0;
            for (int q = 0; q < nbp; ++q) {
                const double dq = __shfl_sync(0xffffffffu, b, q) * rd[q];
                if (lane == q) a.D[(size_t)jt * 32 + q] = dq;
                if (lane > q && lane < nbp) b -= dq * tpp[lane][q];
            }
        }
        deim_grid_sync(a.bar, nblk);
        const double *Wp = R;
        for (int jt0 = 0; jt0 < ntrail; jt0 += 64) {
            const int nc = min(64, ntrail - jt0);
            __syncthreads();
            for (int q = tid; q < 64 * NB; q += THREADS) {
                const int c = q / NB, l = q - c * NB;
                Dt[c][l] = (c < nc && l < nbp) ? __ldcg(a.D + (size_t)(jt0 + c) * 32 + l) : 0.0;
            }
            __syncthreads();
            double *Wt = W + (size_t)(j0 + nbp + jt0) * ld;
            if (VEC) {
                for (int64_t r = r0 + 2 * tid; r < r1; r += RSTEP) {
                    double2 pv[NB];
#pragma unroll
                    for (int l = 0; l < NB; ++l) pv[l] = l < nbp ? *reinterpret_cast<const double2 *>(Wp + (size_t)l * ld + r) : make_double2(0.0, 0.0);
                    for (int c0 = 0; c0 < nc; c0 += 4) {   // four trailing columns at a time: four loads in flight per thread
                        double2 acc[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            acc[u] = c0 + u < nc ? *reinterpret_cast<const double2 *>(Wt + (size_t)(c0 + u) * ld + r) : make_double2(0.0, 0.0);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int c = min(c0 + u, 63);
#pragma unroll
                            for (int l = 0; l < NB; ++l) {
                                acc[u].x = fma(-Dt[c][l], pv[l].x, acc[u].x);
                                acc[u].y = fma(-Dt[c][l], pv[l].y, acc[u].y);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (c0 + u < nc) *reinterpret_cast<double2 *>(Wt + (size_t)(c0 + u) * ld + r) = acc[u];
                    }
                }
            } else {
                for (int64_t r = r0 + tid; r < r1; r += RSTEP) {
                    double pv[NB];
#pragma unroll
                    for (int l = 0; l < NB; ++l) pv[l] = l < nbp ? Wp[(size_t)l * ld + r] : 0.0;
                    for (int c = 0; c < nc; ++c) {
                        double acc = Wt[(size_t)c * ld + r];
#pragma unroll
                        for (int l = 0; l < NB; ++l) acc = fma(-Dt[c][l], pv[l], acc);
                        Wt[(size_t)c * ld + r] = acc;
                    }
                }
            }
        }
        deim_grid_sync(a.bar, nblk);
    }
}

template <int NB, bool VEC, int THREADS> int launch_deim_blocked(rbm_ctx *ctx, DeimArgs &a, int grid)
{
    void *args[] = {&a};
    ProfScope ps(ctx, "k_deim_blocked");
    RBM_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)k_deim_blocked<NB, VEC, THREADS>, dim3(grid), dim3(THREADS), args, 0, ctx->stream));
    RBM_LAUNCH_CHECK(ctx);
    return RBM_OK;
}

// single rank: the whole selection in one cooperative launch
int deim_single_rank(rbm_ctx *ctx, const double *Psi, int64_t n, int64_t ld, int k, int64_t *idx)
{
    DevBuf W, Rb, cand_val, cand_idx, D, d_idx, piv, bar;
    RBM_CUDA(ctx, W.alloc(sizeof(double) * (size_t)n * k));
    RBM_CUDA(ctx, Rb.alloc(sizeof(double) * (size_t)n * std::min(k, 32)));
    RBM_CUDA(ctx, cudaMemcpy2DAsync(W.p, sizeof(double) * n, Psi, sizeof(double) * ld, sizeof(double) * n, k, cudaMemcpyDefault, ctx->stream));
    // one CTA per SM at most; few rows: one CTA per 64 rows (a smaller grid makes the barrier cheaper)
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ctx->sm_count, (n + 63) / 64));
    int64_t rpc = (n + grid - 1) / grid;
    rpc += rpc & 1;
    RBM_CUDA(ctx, cand_val.alloc(sizeof(double) * 2 * grid));
    RBM_CUDA(ctx, cand_idx.alloc(sizeof(int64_t) * 2 * grid));
    RBM_CUDA(ctx, D.alloc(sizeof(double) * 32 * (size_t)k));
    RBM_CUDA(ctx, d_idx.alloc(sizeof(int64_t) * k));
    RBM_CUDA(ctx, piv.alloc(sizeof(double) * k));
    RBM_CUDA(ctx, bar.alloc(sizeof(unsigned) * 2));
    RBM_CUDA(ctx, cudaMemsetAsync(bar.p, 0, sizeof(unsigned) * 2, ctx->stream));
    RBM_CUDA(ctx, cudaMemsetAsync(piv.p, 0, sizeof(double) * k, ctx->stream));
    DeimArgs a{};
    a.W = W.as<double>(); a.R = Rb.as<double>(); a.n = n; a.k = k; a.rows_per_cta = rpc;
    a.cand_val = cand_val.as<double>(); a.cand_idx = cand_idx.as<int64_t>(); a.D = D.as<double>();
    a.idx = d_idx.as<int64_t>(); a.piv = piv.as<double>(); a.bar = bar.as<unsigned>();
    const bool vec = (n % 2 == 0);
    // panel width: 32 while the matrix lives in L2 (fewest barriers per column); 16 when it streams from HBM (fewest bytes)
    const bool narrow = (double)n * k * 8.0 > 96e6;
    if (narrow) RBM_TRY(vec ? (launch_deim_blocked<16, true, 512>(ctx, a, grid)) : (launch_deim_blocked<16, false, 512>(ctx, a, grid)));
    else RBM_TRY(vec ? (launch_deim_blocked<32, true, 256>(ctx, a, grid)) : (launch_deim_blocked<32, false, 256>(ctx, a, grid)));
    std::vector<double> hp(k);
    RBM_CUDA(ctx, cudaMemcpyAsync(idx, d_idx.p, sizeof(int64_t) * k, cudaMemcpyDefault, ctx->stream));
    RBM_CUDA(ctx, cudaMemcpyAsync(hp.data(), piv.p, sizeof(double) * k, cudaMemcpyDeviceToHost, ctx->stream));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // a zero pivot means Psi's columns are linearly dependent at the selected rows
    for (int i = 0; i < k; ++i)
        if (!(std::fabs(hp[i]) > 0.0) || !std::isfinite(hp[i]))
            return rbm_fail(ctx, RBM_ERR_NUMERIC, "rbm_deim: singular interpolation matrix at point %d (SingularException)", i);
    return RBM_OK;
}

}  // namespace

extern "C" int rbm_deim(rbm_ctx *ctx, const double *Psi, int64_t n, int64_t ld, int k, int64_t *idx)
{
    if (!ctx) return RBM_ERR_INVALID;
    if (!Psi || !idx || n < 1 || k < 1 || ld < n) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_deim: bad argument (DimensionMismatch)");
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    // global row offset of this rank's rows
    int64_t row0 = 0, ntotal = n;
    if (ctx->nranks > 1) {
        DevBuf cnt;
        RBM_CUDA(ctx, cnt.alloc(sizeof(double) * ctx->nranks));
        std::vector<double> h(ctx->nranks, 0.0);
        h[ctx->rank] = (double)n;
        RBM_CUDA(ctx, cudaMemcpyAsync(cnt.p, h.data(), sizeof(double) * ctx->nranks, cudaMemcpyHostToDevice, ctx->stream));
        RBM_TRY(rbm_allreduce_sum_f64(ctx, cnt.as<double>(), ctx->nranks));
        RBM_CUDA(ctx, cudaMemcpyAsync(h.data(), cnt.p, sizeof(double) * ctx->nranks, cudaMemcpyDeviceToHost, ctx->stream));
        RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ntotal = 0;
        for (int r = 0; r < ctx->nranks; ++r) {
            if (r < ctx->rank) row0 += (int64_t)h[r];
            ntotal += (int64_t)h[r];
        }
    }
    if (k > ntotal) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_deim: k = %d exceeds the %lld rows", k, (long long)ntotal);
    {
        const char *e = getenv("RBM_DEIM_BLOCKED");
        if (ctx->nranks <= 1 && !(e && e[0] == '0')) return deim_single_rank(ctx, Psi, n, ld, k, idx);
    }
    // working copy of Psi (dense n x k)
    DevBuf W, T, rdiag, d, bvec, rowvals, pair, blk_val, blk_idx, d_idx;
    RBM_CUDA(ctx, W.alloc(sizeof(double) * (size_t)n * k));
    RBM_CUDA(ctx, cudaMemcpy2DAsync(W.p, sizeof(double) * n, Psi, sizeof(double) * ld, sizeof(double) * n, k, cudaMemcpyDefault, ctx->stream));
    // few rows (the reference's sizes): 8 warps split the columns of one 64-row slab so that all SMs get work;
    // many rows: one slab per warp, no block-level combine
    const int64_t slabs = (n + DR_ROWS - 1) / DR_ROWS;
    const bool split = slabs < (int64_t)ctx->sm_count * 32;
    const int nblk = (int)std::min<int64_t>(split ? slabs : (slabs + 7) / 8, (int64_t)ctx->sm_count * 8);
    const int wvec = (n % 2 == 0) ? 1 : 0;   // W is a dense n x k copy: columns are 16-byte aligned iff n is even
    RBM_CUDA(ctx, T.alloc(sizeof(double) * (size_t)k * k));
    RBM_CUDA(ctx, cudaMemsetAsync(T.p, 0, T.bytes, ctx->stream));
    RBM_CUDA(ctx, d.alloc(sizeof(double) * k));
    RBM_CUDA(ctx, rdiag.alloc(sizeof(double) * k));
    RBM_CUDA(ctx, bvec.alloc(sizeof(double) * k));
    RBM_CUDA(ctx, rowvals.alloc(sizeof(double) * k));
    RBM_CUDA(ctx, pair.alloc(sizeof(double) * 2));
    RBM_CUDA(ctx, blk_val.alloc(sizeof(double) * nblk));
    RBM_CUDA(ctx, blk_idx.alloc(sizeof(int64_t) * nblk));
    RBM_CUDA(ctx, d_idx.alloc(sizeof(int64_t) * k));
    RBM_CUDA(ctx, cudaFuncSetAttribute(k_deim_residual<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * k)));
    RBM_CUDA(ctx, cudaFuncSetAttribute(k_deim_residual<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * k)));
    const bool solo = ctx->nranks <= 1;   // no exchange steps: gather, pick and commit are fused into their neighbours
    for (int i = 0; i < k; ++i) {
        if (i > 0) {
            if (!solo) {
                // b = Psi_i at the selected rows (summed over ranks: each row is owned by one rank)
                k_deim_gather_rows<<<(i + 127) / 128, 128, 0, ctx->stream>>>(W.as<double>(), n, d_idx.as<int64_t>(), i, row0, n, i, bvec.as<double>());
                RBM_LAUNCH_CHECK(ctx);
                RBM_TRY(rbm_allreduce_sum_f64(ctx, bvec.as<double>(), i));
            }
            {
                ProfScope ps(ctx, "k_deim_coeffs");
                k_deim_coeffs<<<1, 256, sizeof(double) * i, ctx->stream>>>(T.as<double>(), rdiag.as<double>(), k, i, d.as<double>(), bvec.as<double>(),
                                                                           solo ? W.as<double>() + (size_t)i * n : nullptr, d_idx.as<int64_t>());
            }
            RBM_LAUNCH_CHECK(ctx);
        }
        {
            ProfScope ps(ctx, "k_deim_residual");
            if (split) k_deim_residual<true><<<nblk, 256, sizeof(double) * std::max(i, 1), ctx->stream>>>(W.as<double>(), n, n, i, d.as<double>(), blk_val.as<double>(), blk_idx.as<int64_t>(), wvec);
            else k_deim_residual<false><<<nblk, 256, sizeof(double) * std::max(i, 1), ctx->stream>>>(W.as<double>(), n, n, i, d.as<double>(), blk_val.as<double>(), blk_idx.as<int64_t>(), wvec);
        }
        RBM_LAUNCH_CHECK(ctx);
        if (solo) {
            ProfScope ps(ctx, "k_deim_pick");
            k_deim_pick_commit<<<1, 256, 0, ctx->stream>>>(blk_val.as<double>(), blk_idx.as<int64_t>(), nblk, W.as<double>(), n, i, d_idx.as<int64_t>(),
                                                           T.as<double>(), rdiag.as<double>(), k);
            RBM_LAUNCH_CHECK(ctx);
            continue;
        }
        k_deim_pick<<<1, 256, 0, ctx->stream>>>(blk_val.as<double>(), blk_idx.as<int64_t>(), nblk, row0, pair.as<double>());
        RBM_LAUNCH_CHECK(ctx);
        RBM_TRY(rbm_allreduce_argmax(ctx, pair.as<double>()));
        // row i of the triangular table: the new residual columns 0..i at the new row
        k_deim_gather_cols<<<(i + 1 + 127) / 128, 128, 0, ctx->stream>>>(W.as<double>(), n, pair.as<double>(), row0, n, i + 1, rowvals.as<double>());
        RBM_LAUNCH_CHECK(ctx);
        RBM_TRY(rbm_allreduce_sum_f64(ctx, rowvals.as<double>(), i + 1));
        k_deim_commit<<<(i + 1 + 127) / 128, 128, 0, ctx->stream>>>(pair.as<double>(), rowvals.as<double>(), i, d_idx.as<int64_t>(), T.as<double>(),
                                                                    rdiag.as<double>(), k);
        RBM_LAUNCH_CHECK(ctx);
    }
    RBM_CUDA(ctx, cudaMemcpyAsync(idx, d_idx.p, sizeof(int64_t) * k, cudaMemcpyDefault, ctx->stream));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // a zero pivot means Psi's columns are linearly dependent at the selected rows
    std::vector<double> hT((size_t)k * k);
    RBM_CUDA(ctx, cudaMemcpy(hT.data(), T.p, sizeof(double) * (size_t)k * k, cudaMemcpyDeviceToHost));
    for (int i = 0; i < k; ++i)
        if (!(std::fabs(hT[(size_t)i * k + i]) > 0.0) || !std::isfinite(hT[(size_t)i * k + i]))
            return rbm_fail(ctx, RBM_ERR_NUMERIC, "rbm_deim: singular interpolation matrix at point %d (SingularException)", i);
    return RBM_OK;
}
