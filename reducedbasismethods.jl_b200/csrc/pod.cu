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
    if (ctx->evd.eigvec) cudaFree(ctx->evd.eigvec);
    ctx->evd.eigvec = nullptr;
    ctx->evd.valid = false;
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

// G (device, C x C) = S'S summed over ranks
int gram_device(rbm_ctx *ctx, const ColumnSet &cs, int64_t nrows, double *G)
{
    const int64_t C = cs.ncols;
    RBM_TRY(run_gemm(ctx, true, cs.dev(), cs.dev(), C, C, nrows, cs.vec_ok, true, G, C));
    RBM_TRY(rbm_allreduce_sum_f64(ctx, G, (size_t)C * C));
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
    RBM_CUDA(ctx, out.alloc(sizeof(double *) * (size_t)ncols));
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
    const int64_t C = cs.ncols;
    if (k > C) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pod_evd: k = %d exceeds the %lld snapshot columns", k, (long long)C);
    if (Psi && ldpsi < nrows) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pod_evd: ldpsi < nrows");
    if (C > 32768) return rbm_fail(ctx, RBM_ERR_UNSUPPORTED, "rbm_pod_evd: %lld columns (limit 32768)", (long long)C);

    // ---- Gram matrix (evd.jl:33 / :65) and eigen(G) (evd.jl:35 / :67): symmetric path, eigenvectors
    // overwrite G.  A rank query (Psi == NULL) leaves the decomposition in the context for the next call
    // on the same blocks, which then only forms Psi.
    DevBuf G;
    std::vector<double> lam(C);
    bool from_cache = false;
    {
        rbm_ctx::EvdCache &ec = ctx->evd;
        bool same = ec.valid && ec.nrows == nrows && (int)ec.blocks.size() == nblocks;
        for (int b = 0; same && b < nblocks; ++b) same = ec.blocks[b] == (const void *)blocks[b] && ec.ncols[b] == ncols[b];
        if (same && Psi) {
            G.p = ec.eigvec; G.bytes = sizeof(double) * (size_t)C * C;   // take ownership
            ec.eigvec = nullptr;
            lam = ec.lam;
            from_cache = true;
        } else if (ec.eigvec) {
            cudaFree(ec.eigvec);
            ec.eigvec = nullptr;
        }
        ec.valid = false;
    }
    if (!from_cache) {
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
    }

    // ---- sorteigen: by |lambda| descending, stable (src/eigen.jl:4-7)
    std::vector<int> perm(C);
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return std::fabs(lam[a]) > std::fabs(lam[b]); });
    std::vector<double> Ls(C);
    for (int64_t i = 0; i < C; ++i) Ls[i] = lam[perm[i]];
    // ---- rank from the energy criterion (evd.jl:38-46 / :70-78)
    int kk = k;
    if (kk == 0) {
        double E = 0.0;
        for (int64_t i = 0; i < C; ++i) E += Ls[i];
        double Er = 0.0;
        kk = 1;
        while (kk < C && 1.0 - tolerance > (Er + Ls[kk - 1]) / E) {
            Er += Ls[kk - 1];
            ++kk;
        }
    }
    *k_out = kk;
    if (Lambda) RBM_CUDA(ctx, cudaMemcpyAsync(Lambda, Ls.data(), sizeof(double) * C, cudaMemcpyDefault, ctx->stream));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (!Psi) {   // rank query: keep the decomposition for the call that forms Psi
        rbm_ctx::EvdCache &ec = ctx->evd;
        ec.blocks.assign(blocks, blocks + nblocks);
        ec.ncols.assign(ncols, ncols + nblocks);
        ec.nrows = nrows;
        ec.lam = lam;
        ec.eigvec = G.p;
        G.p = nullptr;   // ownership moves to the cache
        G.bytes = 0;
        ec.valid = true;
        return RBM_OK;
    }
    if (k == 0 && kk > psi_capacity)
        return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pod_evd: rank %d exceeds psi_capacity %d (query with Psi == NULL first)", kk, psi_capacity);

    // ---- Psi = S * Omega[:,1:k] * diag(|Lambda|^-1/2)  (evd.jl:49 / :81).  The sign of each
    // eigenvector is solver-dependent; it is fixed here by making its largest-magnitude entry positive.
    DevBuf dPerm, dScale, dB, dBcols;
    RBM_CUDA(ctx, dPerm.alloc(sizeof(int) * kk));
    RBM_CUDA(ctx, dScale.alloc(sizeof(double) * kk));
    RBM_CUDA(ctx, dB.alloc(sizeof(double) * (size_t)C * kk));
    RBM_CUDA(ctx, cudaMemcpyAsync(dPerm.p, perm.data(), sizeof(int) * kk, cudaMemcpyHostToDevice, ctx->stream));
    k_col_absmax_sign<<<kk, 256, 0, ctx->stream>>>(G.as<double>(), C, dPerm.as<int>(), kk, dScale.as<double>());
    RBM_LAUNCH_CHECK(ctx);
    std::vector<double> sc(kk);
    RBM_CUDA(ctx, cudaMemcpyAsync(sc.data(), dScale.p, sizeof(double) * kk, cudaMemcpyDeviceToHost, ctx->stream));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int j = 0; j < kk; ++j) sc[j] *= 1.0 / std::sqrt(std::fabs(Ls[j]));
    RBM_CUDA(ctx, cudaMemcpyAsync(dScale.p, sc.data(), sizeof(double) * kk, cudaMemcpyHostToDevice, ctx->stream));
    {
        const int64_t n = C * kk;
        k_scale_cols<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(G.as<double>(), C, dPerm.as<int>(),
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

// d = forward substitution:  T d = b,  T[a][j] = W[idx_a, j] (j <= a), b[a] = W[idx_a, i]
__global__ void k_deim_coeffs(const double *W, int64_t ld, const int64_t *idx_local, const double *T, int ldt, int i,
                              double *d, const double *bvec)
{
    // single block; T is the k x k lower-triangular table (row a = selected row a), bvec = Psi_i at the selected rows
    extern __shared__ double sh[];
    double *b = sh;  // [i]
    for (int a = threadIdx.x; a < i; a += blockDim.x) b[a] = bvec[a];
    __syncthreads();
    __shared__ double dj;
    for (int j = 0; j < i; ++j) {
        if (threadIdx.x == 0) {
            dj = b[j] / T[(size_t)j * ldt + j];
            d[j] = dj;
        }
        __syncthreads();
        for (int a = j + 1 + threadIdx.x; a < i; a += blockDim.x) b[a] -= dj * T[(size_t)a * ldt + j];
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

// r = W[:,i] - W[:,0:i] d  (in place), block-level signed arg-max (first index on ties)
__global__ void __launch_bounds__(256) k_deim_residual(double *W, int64_t ld, int64_t n, int i, const double *d,
                                                      double *blk_val, int64_t *blk_idx)
{
    extern __shared__ double sd[];  // d[i]
    for (int j = threadIdx.x; j < i; j += blockDim.x) sd[j] = d[j];
    __syncthreads();
    double bv = -INFINITY;
    int64_t bi = INT64_MAX;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        double acc = W[(size_t)i * ld + r];
        for (int j = 0; j < i; ++j) acc = fma(-sd[j], W[(size_t)j * ld + r], acc);
        if (i > 0) W[(size_t)i * ld + r] = acc;
        if (acc > bv) { bv = acc; bi = r; }
    }
    __shared__ double sv[256];
    __shared__ int64_t si[256];
    sv[threadIdx.x] = bv; si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            double v2 = sv[threadIdx.x + o]; int64_t i2 = si[threadIdx.x + o];
            if (v2 > sv[threadIdx.x] || (v2 == sv[threadIdx.x] && i2 < si[threadIdx.x])) { sv[threadIdx.x] = v2; si[threadIdx.x] = i2; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { blk_val[blockIdx.x] = sv[0]; blk_idx[blockIdx.x] = si[0]; }
}

__global__ void k_deim_pick(const double *blk_val, const int64_t *blk_idx, int nblk, int64_t row0, double *pair)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double bv = -INFINITY;
        int64_t bi = INT64_MAX;
        for (int b = 0; b < nblk; ++b)
            if (blk_val[b] > bv || (blk_val[b] == bv && blk_idx[b] < bi)) { bv = blk_val[b]; bi = blk_idx[b]; }
        pair[0] = bv;
        pair[1] = (double)(bi + row0);  // global row (exact for rows < 2^53)
    }
}

__global__ void k_deim_commit(const double *pair, const double *rowvals, int i, int64_t *idx_global, double *T, int ldt)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) idx_global[i] = (int64_t)pair[1];
    if (j <= i) T[(size_t)i * ldt + j] = rowvals[j];
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
    // working copy of Psi (dense n x k)
    DevBuf W, T, d, bvec, rowvals, pair, blk_val, blk_idx, d_idx;
    RBM_CUDA(ctx, W.alloc(sizeof(double) * (size_t)n * k));
    RBM_CUDA(ctx, cudaMemcpy2DAsync(W.p, sizeof(double) * n, Psi, sizeof(double) * ld, sizeof(double) * n, k, cudaMemcpyDefault, ctx->stream));
    const int nblk = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 4);
    RBM_CUDA(ctx, T.alloc(sizeof(double) * (size_t)k * k));
    RBM_CUDA(ctx, cudaMemsetAsync(T.p, 0, T.bytes, ctx->stream));
    RBM_CUDA(ctx, d.alloc(sizeof(double) * k));
    RBM_CUDA(ctx, bvec.alloc(sizeof(double) * k));
    RBM_CUDA(ctx, rowvals.alloc(sizeof(double) * k));
    RBM_CUDA(ctx, pair.alloc(sizeof(double) * 2));
    RBM_CUDA(ctx, blk_val.alloc(sizeof(double) * nblk));
    RBM_CUDA(ctx, blk_idx.alloc(sizeof(int64_t) * nblk));
    RBM_CUDA(ctx, d_idx.alloc(sizeof(int64_t) * k));
    RBM_CUDA(ctx, cudaFuncSetAttribute(k_deim_residual, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * k)));
    for (int i = 0; i < k; ++i) {
        if (i > 0) {
            // b = Psi_i at the selected rows (summed over ranks: each row is owned by one rank)
            k_deim_gather_rows<<<(i + 127) / 128, 128, 0, ctx->stream>>>(W.as<double>(), n, d_idx.as<int64_t>(), i, row0, n, i, bvec.as<double>());
            RBM_LAUNCH_CHECK(ctx);
            RBM_TRY(rbm_allreduce_sum_f64(ctx, bvec.as<double>(), i));
            k_deim_coeffs<<<1, 256, sizeof(double) * i, ctx->stream>>>(W.as<double>(), n, d_idx.as<int64_t>(), T.as<double>(), k, i, d.as<double>(), bvec.as<double>());
            RBM_LAUNCH_CHECK(ctx);
        }
        k_deim_residual<<<nblk, 256, sizeof(double) * std::max(i, 1), ctx->stream>>>(W.as<double>(), n, n, i, d.as<double>(), blk_val.as<double>(), blk_idx.as<int64_t>());
        RBM_LAUNCH_CHECK(ctx);
        k_deim_pick<<<1, 32, 0, ctx->stream>>>(blk_val.as<double>(), blk_idx.as<int64_t>(), nblk, row0, pair.as<double>());
        RBM_LAUNCH_CHECK(ctx);
        RBM_TRY(rbm_allreduce_argmax(ctx, pair.as<double>()));
        // row i of the triangular table: the new residual columns 0..i at the new row
        k_deim_gather_cols<<<(i + 1 + 127) / 128, 128, 0, ctx->stream>>>(W.as<double>(), n, pair.as<double>(), row0, n, i + 1, rowvals.as<double>());
        RBM_LAUNCH_CHECK(ctx);
        RBM_TRY(rbm_allreduce_sum_f64(ctx, rowvals.as<double>(), i + 1));
        k_deim_commit<<<(i + 1 + 127) / 128, 128, 0, ctx->stream>>>(pair.as<double>(), rowvals.as<double>(), i, d_idx.as<int64_t>(), T.as<double>(), k);
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
