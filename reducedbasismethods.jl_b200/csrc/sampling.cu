// sampling.cu -- the two small host-side stages around the PIC solve, moved to the device so that a
// parameter sweep never leaves HBM:
//   rbm_sample_bump_on_tail   the accept-reject initial-condition draw that
//                             scripts/bump_on_tail_1_training.jl:76-79 gets from VlasovMethods.BumpOnTail
//                             (external package), re-designed around a counter-based generator
//   rbm_growth_rate           find_maxima + get_regression_alpha_beta of src/regression.jl:4-24
//
// The sampler is NOT stream-compatible with Julia's MersenneTwister (nothing on a GPU can be): the
// draw of particle p depends only on (seed, p) through Philox4x32-10, so it is independent of the launch
// geometry and of the number of ranks, and oracle/sampler_numpy.py restates it for the parity test.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace rbm {

struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4]) const
    {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            c0 = hi1 ^ c1 ^ a;
            c1 = lo1;
            c2 = hi0 ^ c3 ^ b;
            c3 = lo0;
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};

// two uniforms in [0, 1) with 53 random bits each from one Philox block
__device__ __forceinline__ void uniforms(const uint32_t (&r)[4], double &u0, double &u1)
{
    u0 = (double)((((uint64_t)r[0] << 32) | r[1]) >> 11) * 0x1.0p-53;
    u1 = (double)((((uint64_t)r[2] << 32) | r[3]) >> 11) * 0x1.0p-53;
}

struct SampleArgs {
    int64_t n, first;            // particles of this call; global index of the first one (particle-chunk ranks)
    uint32_t k0, k1;
    double L, kappa, eps, a, v0, sigma, w;
    double *x, *v, *wout;
    unsigned long long *rejected;   // diagnostics: total rejected candidates
};

// f(x, v) = (1 + eps cos(kappa x)) [(1-a) N(0,1)(v) + a N(v0, sigma^2)(v)] on [0, L):
// v exactly from the mixture (Box-Muller), x by accept-reject under the envelope 1 + eps.
__global__ void __launch_bounds__(256) k_sample_bump_on_tail(const SampleArgs s)
{
    const Philox ph{s.k0, s.k1};
    unsigned long long rej = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < s.n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t p = (uint64_t)(s.first + i);
        const uint32_t p0 = (uint32_t)p, p1 = (uint32_t)(p >> 32);
        uint32_t r[4];
        double u0, u1, u2, u3;
        ph(p0, p1, 0u, 0u, r);
        uniforms(r, u0, u1);
        ph(p0, p1, 1u, 0u, r);
        uniforms(r, u2, u3);
        const double z = sqrt(-2.0 * log(1.0 - u2)) * cos(6.283185307179586 * u3);
        s.v[i] = u0 < s.a ? s.v0 + s.sigma * z : z;
        double cand = 0.0;
        for (uint32_t j = 2;; ++j) {   // acceptance probability >= (1-eps)/(1+eps): terminates almost surely
            ph(p0, p1, j, 0u, r);
            double uc, ua;
            uniforms(r, uc, ua);
            cand = uc * s.L;
            if (ua * (1.0 + s.eps) <= 1.0 + s.eps * cos(s.kappa * cand) || j == 0xFFFFFFFFu) break;
            ++rej;
        }
        s.x[i] = cand;
        if (s.wout) s.wout[i] = s.w;
    }
    if (s.rejected && rej) atomicAdd(s.rejected, rej);
}

// One block per parameter sample.  W is (ns x nparam) column-major with leading dimension ldw.
// find_maxima(w, n): indices i in n .. ns-n-1 (0-based) with w[i-n..i] non-decreasing and w[i..i+n]
// non-increasing (issorted with isless, regression.jl:6-7); m = maxima[inds]; beta = cov(t[m], log W[m]) /
// var(t[m]) (uncorrected), alpha = mean(log W[m]) - beta mean(t[m]).
__global__ void __launch_bounds__(256) k_growth_rate(const double *t, const double *W, int ns, int64_t ldw, int n,
                                                    const int *inds, int ninds, double *alpha, double *beta,
                                                    int *status, int *found)
{
    extern __shared__ int sel[];   // [ninds] selected maxima
    const double *w = W + (size_t)blockIdx.x * ldw;
    // the maxima list is short and must be in increasing order: a single thread walks the trace,
    // the window tests of one candidate are cheap (2n comparisons)
    if (threadIdx.x == 0) {
        int cnt = 0, maxind = -1;
        for (int q = 0; q < ninds; ++q) maxind = max(maxind, inds[q]);
        for (int i = n; i < ns - n; ++i) {
            bool up = true, down = true;
            for (int q = i - n; q < i && up; ++q) up = !(w[q + 1] < w[q]);          // issorted: no descent
            for (int q = i; q < i + n && up && down; ++q) down = !(w[q] < w[q + 1]);  // rev: no ascent
            if (up && down) {
                for (int q = 0; q < ninds; ++q)
                    if (inds[q] == cnt) sel[q] = i;
                ++cnt;
                if (cnt > maxind) break;
            }
        }
        if (found) found[blockIdx.x] = cnt;
        if (cnt <= maxind) {
            atomicExch(status, 1);   // BoundsError in the reference: fewer maxima than inds asks for
            alpha[blockIdx.x] = beta[blockIdx.x] = nan("");
        } else {
            double mt = 0.0, ml = 0.0;
            for (int q = 0; q < ninds; ++q) {
                mt += t[sel[q]];
                ml += log(w[sel[q]]);
            }
            mt /= ninds;
            ml /= ninds;
            double cov = 0.0, var = 0.0;
            for (int q = 0; q < ninds; ++q) {
                const double dt = t[sel[q]] - mt;
                cov += dt * (log(w[sel[q]]) - ml);
                var += dt * dt;
            }
            cov /= ninds;
            var /= ninds;
            const double b = cov / var;
            beta[blockIdx.x] = b;
            alpha[blockIdx.x] = ml - b * mt;
        }
    }
}

}  // namespace rbm

using namespace rbm;

extern "C" int rbm_sample_bump_on_tail(rbm_ctx *ctx, int64_t n, int64_t first, uint64_t seed, double kappa, double eps,
                                       double a, double v0, double sigma, double w, double *x, double *v, double *wout,
                                       int64_t *rejected)
{
    if (!ctx) return RBM_ERR_INVALID;
    if (n < 0 || first < 0 || !x || !v) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_sample_bump_on_tail: bad argument");
    if (!(kappa > 0.0) || !(eps >= 0.0 && eps < 1.0) || !(a >= 0.0 && a <= 1.0) || !(sigma > 0.0))
        return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_sample_bump_on_tail: need kappa > 0, 0 <= eps < 1, 0 <= a <= 1, sigma > 0");
    if (rejected) *rejected = 0;
    if (n == 0) return RBM_OK;
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    DevOut dx, dv, dw;
    DevBuf drej;
    RBM_TRY(dx.prepare(ctx, x, sizeof(double) * (size_t)n));
    RBM_TRY(dv.prepare(ctx, v, sizeof(double) * (size_t)n));
    if (wout) RBM_TRY(dw.prepare(ctx, wout, sizeof(double) * (size_t)n));
    RBM_CUDA(ctx, drej.alloc(sizeof(unsigned long long)));
    RBM_CUDA(ctx, cudaMemsetAsync(drej.p, 0, sizeof(unsigned long long), ctx->stream));
    SampleArgs s{};
    s.n = n; s.first = first;
    s.k0 = (uint32_t)seed; s.k1 = (uint32_t)(seed >> 32);
    s.L = 2.0 * 3.14159265358979323846 / kappa;
    s.kappa = kappa; s.eps = eps; s.a = a; s.v0 = v0; s.sigma = sigma; s.w = w;
    s.x = dx.as<double>(); s.v = dv.as<double>(); s.wout = wout ? dw.as<double>() : nullptr;
    s.rejected = drej.as<unsigned long long>();
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8);
    {
        ProfScope ps(ctx, "k_sample_bump_on_tail");
        k_sample_bump_on_tail<<<grid, 256, 0, ctx->stream>>>(s);
        RBM_LAUNCH_CHECK(ctx);
    }
    RBM_TRY(dx.flush(ctx));
    RBM_TRY(dv.flush(ctx));
    if (wout) RBM_TRY(dw.flush(ctx));
    unsigned long long hrej = 0;
    if (rejected) RBM_CUDA(ctx, cudaMemcpyAsync(&hrej, drej.p, sizeof(hrej), cudaMemcpyDeviceToHost, ctx->stream));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (rejected) *rejected = (int64_t)hrej;
    return RBM_OK;
}

extern "C" int rbm_growth_rate(rbm_ctx *ctx, const double *t, const double *W, int64_t ns, int64_t nparam, int64_t ldw,
                               int n, const int *inds, int ninds, double *alpha, double *beta, int *nmaxima)
{
    if (!ctx) return RBM_ERR_INVALID;
    if (!t || !W || !inds || !alpha || !beta || ns < 1 || nparam < 0 || ldw < ns || n < 0 || ninds < 1 || ns > (1 << 30))
        return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_growth_rate: bad argument");
    for (int q = 0; q < ninds; ++q)
        if (inds[q] < 0) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_growth_rate: inds are 0-based positions in the list of maxima");
    if (nparam == 0) return RBM_OK;
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    DevIn dt, dW, di;
    DevOut da, db, dn;
    DevBuf dstat;
    RBM_TRY(dt.stage(ctx, t, sizeof(double) * (size_t)ns));
    RBM_TRY(dW.stage(ctx, W, sizeof(double) * (size_t)ldw * (size_t)nparam));
    RBM_TRY(di.stage(ctx, inds, sizeof(int) * (size_t)ninds));
    RBM_TRY(da.prepare(ctx, alpha, sizeof(double) * (size_t)nparam));
    RBM_TRY(db.prepare(ctx, beta, sizeof(double) * (size_t)nparam));
    if (nmaxima) RBM_TRY(dn.prepare(ctx, nmaxima, sizeof(int) * (size_t)nparam));
    RBM_CUDA(ctx, dstat.alloc(sizeof(int)));
    RBM_CUDA(ctx, cudaMemsetAsync(dstat.p, 0, sizeof(int), ctx->stream));
    {
        ProfScope ps(ctx, "k_growth_rate");
        k_growth_rate<<<(unsigned)nparam, 32, sizeof(int) * (size_t)ninds, ctx->stream>>>(
            dt.as<double>(), dW.as<double>(), (int)ns, ldw, n, di.as<int>(), ninds, da.as<double>(), db.as<double>(),
            dstat.as<int>(), nmaxima ? dn.as<int>() : nullptr);
        RBM_LAUNCH_CHECK(ctx);
    }
    RBM_TRY(da.flush(ctx));
    RBM_TRY(db.flush(ctx));
    if (nmaxima) RBM_TRY(dn.flush(ctx));
    int hstat = 0;
    RBM_CUDA(ctx, cudaMemcpyAsync(&hstat, dstat.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (hstat) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_growth_rate: a trace has fewer local maxima than inds selects (BoundsError)");
    return RBM_OK;
}
