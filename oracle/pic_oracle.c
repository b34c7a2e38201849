/*
 * oracle/pic_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, FP64, no FMA contraction) of the full-order 1D1V
 * Vlasov-Poisson particle-in-cell path behind the bump-on-tail training set of
 * JuliaRCM/ReducedBasisMethods.jl.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * PARITY UNPINNED for this half of the path: the arithmetic lives in the
 * un-vendored registry dependencies VlasovMethods.jl (compat "0.1"),
 * PoissonSolvers.jl ("0.1.0") and ParticleMethods.jl ("0.1.0")
 * (reference Project.toml:19,21,26,36,38,41; no Manifest.toml), and no
 * reference test deposits, solves, gathers or integrates
 * (reference test/trainingset_tests.jl:21-22 only constructs the types).
 * The restatement follows the reference's own call sites and its in-tree twin
 * of the time loop, and SURVEY.md Appendix A for the spline/Poisson maths:
 *
 *   time loop, dt scaling, save cadence .. src/particles/time_marching.jl:119-149
 *   diagnostics W, K, M .................. src/particles/time_marching.jl:95-101
 *   field protocol update!/efield! ....... src/particles/electric_field.jl:22-35
 *   snapshot layout (np fastest) ......... src/particles/snapshots.jl:16-25,
 *                                          scripts/bump_on_tail_1_training.jl:100-108
 *   field energy W = chi^2 * 1/2 phi'K phi notebooks/VP-ROM-Plotting.ipynb cell 24,
 *                                          notebooks/VP-ROM-Training.ipynb cell 9
 *   gauge (K+R) phi = (1-R) rho .......... src/gridbased/poisson.jl:34-43
 *
 * Compile with -ffp-contract=off (see oracle/Makefile): Julia never contracts
 * `x .+= 0.5 .* dt .* v` into an FMA, and the cell index must be bit-exact.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_NH 4096

/* Julia `mod(x::Float64, y::Float64)`: fmod followed by a sign fix-up
 * (Base float.jl); x is never wrapped in storage, periodicity lives in the
 * evaluation (reference src/particles/electric_field.jl:23-24 reconstructs
 * x = Psi*z every step). */
static inline double julia_mod(double x, double y)
{
    double r = fmod(x, y);
    if (r == 0.0) return copysign(r, y);
    if ((r > 0.0) != (y > 0.0)) return r + y;
    return r;
}

/* s = mod(x,L)/h, c = floor(s) clamped to nh-1, xi = s - c. */
static inline void locate1(double x, double L, double h, int nh, int32_t *c, double *xi)
{
    double s = julia_mod(x, L) / h;
    int ci = (int)floor(s);
    if (ci > nh - 1) ci = nh - 1;
    *c = ci;
    *xi = s - (double)ci;
}

/* cubic B-spline values on the four taps j = (c+m) mod nh, m = 0..3 */
static inline void bspline3(double xi, double b[4])
{
    double u = 1.0 - xi;
    b[0] = u * u * u / 6.0;
    b[1] = (3.0 * xi * xi * xi - 6.0 * xi * xi + 4.0) / 6.0;
    b[2] = (-3.0 * xi * xi * xi + 3.0 * xi * xi + 3.0 * xi + 1.0) / 6.0;
    b[3] = xi * xi * xi / 6.0;
}

/* derivatives d/dx of the same four taps */
static inline void bspline3_deriv(double xi, double h, double d[4])
{
    double u = 1.0 - xi;
    d[0] = -(u * u) / 2.0 / h;
    d[1] = (9.0 * xi * xi - 12.0 * xi) / 6.0 / h;
    d[2] = (-9.0 * xi * xi + 6.0 * xi + 3.0) / 6.0 / h;
    d[3] = (xi * xi) / 2.0 / h;
}

void orc_locate(const double *x, int64_t n, double L, int nh, int32_t *cell, double *xi)
{
    double h = L / (double)nh;
    for (int64_t i = 0; i < n; ++i) locate1(x[i], L, h, nh, &cell[i], &xi[i]);
}

/*
 * Charge deposition rhs_j = sum_p w_p B_j(x_p)   (PoissonSolvers
 * `rhs_particles_PBSBasis`, imported at src/particles/time_marching.jl:3).
 * extended != 0 accumulates in long double so that the oracle's own summation
 * roundoff (sqrt(n)*eps, amplified ~30x by the mean removal) stays below the
 * 1e-12 parity bar at 1e7 particles; extended == 0 is the literal
 * `rhs[j] += w*B` in double.
 */
void orc_deposit(const double *x, const double *w, double w_uniform, int64_t n,
                 double L, int nh, int extended, int nthreads, double *rhs)
{
    double h = L / (double)nh;
    if (nthreads < 1) nthreads = 1;
    long double *accl = (long double *)calloc((size_t)nthreads * nh, sizeof(long double));
    double *accd = (double *)calloc((size_t)nthreads * nh, sizeof(double));
#pragma omp parallel num_threads(nthreads)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
        int t = 0, nt = 1;
#endif
        int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
        long double *al = accl + (size_t)t * nh;
        double *ad = accd + (size_t)t * nh;
        for (int64_t i = lo; i < hi; ++i) {
            int32_t c; double xi, b[4];
            locate1(x[i], L, h, nh, &c, &xi);
            bspline3(xi, b);
            double wi = w ? w[i] : w_uniform;
            for (int m = 0; m < 4; ++m) {
                int j = (c + m) % nh;
                if (extended) al[j] += (long double)(wi * b[m]);
                else ad[j] += wi * b[m];
            }
        }
    }
    for (int j = 0; j < nh; ++j) {
        if (extended) {
            long double s = 0.0L;
            for (int t = 0; t < nthreads; ++t) s += accl[(size_t)t * nh + j];
            rhs[j] = (double)s;
        } else {
            double s = 0.0;
            for (int t = 0; t < nthreads; ++t) s += accd[(size_t)t * nh + j];
            rhs[j] = s;
        }
    }
    free(accl); free(accd);
}

/* circulant cubic-spline stiffness row:
 * K_s = (1/h) circ(-1,-24,-15,80,-15,-24,-1)/120 */
static const double KS_BAND[7] = { -1.0, -24.0, -15.0, 80.0, -15.0, -24.0, -1.0 };

void orc_stiffness(int nh, double L, double *K /* nh*nh, row-major (symmetric) */)
{
    double h = L / (double)nh;
    memset(K, 0, sizeof(double) * (size_t)nh * nh);
    for (int i = 0; i < nh; ++i)
        for (int d = -3; d <= 3; ++d) {
            int j = ((i + d) % nh + nh) % nh;
            K[(size_t)i * nh + j] += KS_BAND[d + 3] / 120.0 / h;
        }
}

/* dense LU with partial pivoting (what Julia's `\` does for a square matrix) */
static int lu_solve(double *A, double *b, int n)
{
    for (int k = 0; k < n; ++k) {
        int p = k; double mx = fabs(A[(size_t)k * n + k]);
        for (int i = k + 1; i < n; ++i)
            if (fabs(A[(size_t)i * n + k]) > mx) { mx = fabs(A[(size_t)i * n + k]); p = i; }
        if (mx == 0.0) return -1;
        if (p != k) {
            for (int j = 0; j < n; ++j) { double t = A[(size_t)k*n+j]; A[(size_t)k*n+j] = A[(size_t)p*n+j]; A[(size_t)p*n+j] = t; }
            double t = b[k]; b[k] = b[p]; b[p] = t;
        }
        for (int i = k + 1; i < n; ++i) {
            double f = A[(size_t)i * n + k] / A[(size_t)k * n + k];
            if (f != 0.0) {
                for (int j = k + 1; j < n; ++j) A[(size_t)i*n+j] -= f * A[(size_t)k*n+j];
                b[i] -= f * b[k];
            }
        }
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int j = i + 1; j < n; ++j) s -= A[(size_t)i*n+j] * b[j];
        b[i] = s / A[(size_t)i*n+i];
    }
    return 0;
}

/*
 * Poisson solve of `update!` (PoissonSolverPBSplines.solve! + ScaledPoissonField):
 * remove the mean (neutralising background), solve K_s phi0 = -(rhs - mean) in
 * the zero-mean gauge via (K_s + R) with R = ones/nh (src/gridbased/poisson.jl:34-43),
 * phi = phi0/chi^2, W = chi^2 * 1/2 phi' K_s phi.
 */
int orc_poisson(const double *rhs, int nh, double L, double chi, double *phi, double *Wout)
{
    if (nh > ORC_MAX_NH) return -2;
    double *K = (double *)malloc(sizeof(double) * (size_t)nh * nh);
    double *A = (double *)malloc(sizeof(double) * (size_t)nh * nh);
    double *b = (double *)malloc(sizeof(double) * nh);
    orc_stiffness(nh, L, K);
    double mean = 0.0;
    for (int j = 0; j < nh; ++j) mean += rhs[j];
    mean /= (double)nh;
    for (int j = 0; j < nh; ++j) b[j] = -(rhs[j] - mean);
    for (size_t i = 0; i < (size_t)nh * nh; ++i) A[i] = K[i] + 1.0 / (double)nh;
    int rc = lu_solve(A, b, nh);
    if (rc == 0) {
        for (int j = 0; j < nh; ++j) phi[j] = b[j] / (chi * chi);
        if (Wout) {
            double e = 0.0;
            for (int i = 0; i < nh; ++i) {
                double r = 0.0;
                for (int j = 0; j < nh; ++j) r += K[(size_t)i * nh + j] * phi[j];
                e += phi[i] * r;
            }
            *Wout = chi * chi * 0.5 * e;
        }
    }
    free(K); free(A); free(b);
    return rc;
}

/* Field gather a_p = sum_j phi_j B_j'(x_p)   (PoissonSolvers `eval_deriv_PBSBasis`,
 * imported at src/particles/time_marching.jl:3; `efield!` at electric_field.jl:27-31) */
void orc_gather(const double *phi, const double *x, int64_t n, double L, int nh,
                int nthreads, double *a)
{
    double h = L / (double)nh;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        int32_t c; double xi, d[4];
        locate1(x[i], L, h, nh, &c, &xi);
        bspline3_deriv(xi, h, d);
        double s = 0.0;
        for (int m = 0; m < 4; ++m) s += phi[(c + m) % nh] * d[m];
        a[i] = s;
    }
}

/*
 * integrate_vp! for ONE parameter sample (call site scripts/bump_on_tail_1_training.jl:97;
 * loop structure = src/particles/time_marching.jl:119-149 with Psi = I).
 *   dt_eff = dt*chi (:127); x += 0.5*dt_eff*v (:137); a = E(x) (:140);
 *   v += dt_eff*a (:143); x += 0.5*dt_eff*v (:146); save when it % nsave == 0 (:82).
 * At a save (time_marching.jl:86-101): X,V = current x,v; an update! at the SAVED
 * positions gives Phi and W; K = v'diag(w)v/2; M = w'v.  A at a save is the
 * acceleration used in that step's kick; at the initial save it is the field of
 * the initial update! gathered at x0 (unpinned).
 * x, v (length np) are NOT modified (the reference re-uses `particles` for every
 * sample: scripts/bump_on_tail_1_training.jl:87-97).
 * Outputs are column-major with the particle index fastest: X[i + np*ts].
 */
int orc_integrate(const double *x0, const double *v0, const double *w, double w_uniform,
                  int64_t np, int nh, double L, double chi, double dt, int nt, int ns,
                  double *X, double *V, double *A, double *Phi, double *W, double *K, double *M,
                  int extended, int nthreads)
{
    if (ns < 2 || nt < 1 || nh < 4 || nh > ORC_MAX_NH) return -1;
    int nsave = nt / (ns - 1);
    if (nsave < 1) return -1;
    if (nthreads < 1) nthreads = 1;
    double *x = (double *)malloc(sizeof(double) * np);
    double *v = (double *)malloc(sizeof(double) * np);
    double *a = (double *)malloc(sizeof(double) * np);
    double rhs[ORC_MAX_NH], phi[ORC_MAX_NH];
    memcpy(x, x0, sizeof(double) * np);
    memcpy(v, v0, sizeof(double) * np);
    double dte = dt * chi;
    double hdt = 0.5 * dte;
    int ts = 0;
    for (int it = 0; it <= nt; ++it) {
        if (it > 0) {
#pragma omp parallel for num_threads(nthreads) schedule(static)
            for (int64_t i = 0; i < np; ++i) x[i] = x[i] + hdt * v[i];
            orc_deposit(x, w, w_uniform, np, L, nh, extended, nthreads, rhs);
            if (orc_poisson(rhs, nh, L, chi, phi, NULL)) goto fail;
            orc_gather(phi, x, np, L, nh, nthreads, a);
#pragma omp parallel for num_threads(nthreads) schedule(static)
            for (int64_t i = 0; i < np; ++i) {
                v[i] = v[i] + dte * a[i];
                x[i] = x[i] + hdt * v[i];
            }
        }
        if (it % nsave == 0 && ts < ns) {
            double Wt;
            orc_deposit(x, w, w_uniform, np, L, nh, extended, nthreads, rhs);
            if (orc_poisson(rhs, nh, L, chi, phi, &Wt)) goto fail;
            if (it == 0) orc_gather(phi, x, np, L, nh, nthreads, a);
            if (X) memcpy(X + (size_t)np * ts, x, sizeof(double) * np);
            if (V) memcpy(V + (size_t)np * ts, v, sizeof(double) * np);
            if (A) memcpy(A + (size_t)np * ts, a, sizeof(double) * np);
            if (Phi) memcpy(Phi + (size_t)nh * ts, phi, sizeof(double) * nh);
            long double k = 0.0L, m = 0.0L;
            double kd = 0.0, md = 0.0;
            for (int64_t i = 0; i < np; ++i) {
                double wi = w ? w[i] : w_uniform;
                if (extended) { k += (long double)(v[i] * wi * v[i]); m += (long double)(wi * v[i]); }
                else { kd += v[i] * wi * v[i]; md += wi * v[i]; }
            }
            if (W) W[ts] = Wt;
            if (K) K[ts] = extended ? (double)(k / 2.0L) : kd / 2.0;
            if (M) M[ts] = extended ? (double)m : md;
            ++ts;
        }
    }
    free(x); free(v); free(a);
    return 0;
fail:
    free(x); free(v); free(a);
    return -3;
}

/* One drift-kick-drift step from a given state, in place; returns rhs, phi, a of
 * the step.  Used by the per-step parity tests (identical inputs on both sides). */
int orc_step(double *x, double *v, const double *w, double w_uniform, int64_t np, int nh,
             double L, double chi, double dt, int extended, int nthreads,
             double *rhs, double *phi, double *a)
{
    double dte = dt * chi, hdt = 0.5 * dte;
    for (int64_t i = 0; i < np; ++i) x[i] = x[i] + hdt * v[i];
    orc_deposit(x, w, w_uniform, np, L, nh, extended, nthreads, rhs);
    if (orc_poisson(rhs, nh, L, chi, phi, NULL)) return -3;
    orc_gather(phi, x, np, L, nh, nthreads, a);
    for (int64_t i = 0; i < np; ++i) {
        v[i] = v[i] + dte * a[i];
        x[i] = x[i] + hdt * v[i];
    }
    return 0;
}

/*
 * Check of the division-free device formulation against IEEE division:
 * counts the inputs for which  s = fma(fma(-h, r*y, r), y, r*y)  with y = RN(1/h)
 * (Markstein's correction, used by the CUDA locate) differs from r / h.
 * Deterministic LCG so the test needs no RNG state.
 */
int64_t orc_check_division(double L, int nh, int64_t n, uint64_t seed)
{
    double h = L / (double)nh, y = 1.0 / h;
    int64_t bad = 0;
    uint64_t st = seed * 6364136223846793005ULL + 1442695040888963407ULL;
    for (int64_t i = 0; i < n; ++i) {
        st = st * 6364136223846793005ULL + 1442695040888963407ULL;
        double u = (double)(st >> 11) * (1.0 / 9007199254740992.0);
        double r = u * L;
        if ((i & 7) == 0) { /* cluster some samples at the knots */
            int k = (int)((st >> 3) % (uint64_t)(nh + 1));
            r = nextafter((double)k * h, (st & 1) ? 0.0 : 2.0 * L);
            if (r < 0.0) r = 0.0;
            if (r > L) r = L;
        }
        double q0 = r * y;
        double e = fma(-h, q0, r);
        double s = fma(e, y, q0);
        if (s != r / h) ++bad;
    }
    return bad;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
