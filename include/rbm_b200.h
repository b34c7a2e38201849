/*
 * rbm_b200.h -- C ABI of librbm_b200.so: the B200-native (sm_100a, FP64) replacement
 * for the offline hot path of JuliaRCM/ReducedBasisMethods.jl:
 *   (1) the full-order 1D1V Vlasov-Poisson particle-in-cell solve that fills the
 *       bump-on-tail TrainingSet, and
 *   (2) the snapshot POD (Gram + EVD), basis projection and DEIM point selection.
 *
 * The reference has no FFI layer; its extension points are Julia dispatch
 * interfaces.  Each entry point below names the reference interface it stands
 * behind (paths relative to the reference checkout).  A Julia host reaches these
 * through `ccall`; tests/ reach them through ctypes.
 *
 * Conventions
 *   - every array is column-major FP64, exactly the memory of a Julia Array{Float64};
 *     the library never retains or frees a caller pointer after the call returns;
 *   - every `double*`/`const double*` array argument may be a HOST pointer (pageable or
 *     pinned) or a DEVICE pointer on the context's GPU; the library detects which
 *     (cudaPointerGetAttributes) and stages through device memory when needed;
 *   - every function returns RBM_OK (0) or a negative error code and stores a message
 *     retrievable with rbm_last_error(); no C++ exception crosses the boundary and the
 *     library never aborts.  The Julia wrapper turns non-zero into `error(...)`, the
 *     reference's own error style (`@assert`, `throw(DimensionMismatch())`:
 *     src/parameter.jl:21-24, src/gridbased/poisson.jl:4);
 *   - there is NO CPU fallback: without a usable sm_100 device rbm_init fails;
 *   - calls on one rbm_ctx must be serialised by the caller (the reference is
 *     single-threaded); different contexts (one per GPU) may be driven concurrently.
 *   - indices crossing the boundary are 0-based (the Julia wrapper adds 1).
 */
#ifndef RBM_B200_H
#define RBM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RBM_VERSION 100 /* 0.1.0 */

#define RBM_OK 0
#define RBM_ERR_INVALID (-1)     /* bad argument / dimension mismatch            */
#define RBM_ERR_CUDA (-2)        /* CUDA runtime or driver error                 */
#define RBM_ERR_NOMEM (-3)       /* device or host allocation failed             */
#define RBM_ERR_UNSUPPORTED (-4) /* e.g. spline degree > 3, no sm_100 device     */
#define RBM_ERR_NUMERIC (-5)     /* eigensolver did not converge, singular DEIM  */
#define RBM_ERR_COMM (-6)        /* NCCL not available / communicator failure    */

typedef struct rbm_ctx rbm_ctx; /* one GPU, one stream, optional communicator */
typedef struct rbm_pic rbm_pic; /* particles + spline Poisson field on one GPU */

/* ------------------------------------------------------------------ context */

int rbm_version(void);

/* Bind a context to CUDA device `device`.  Fails with RBM_ERR_UNSUPPORTED unless
 * the device has compute capability 10.x (the kernels are sm_100a only). */
int rbm_init(int device, rbm_ctx **out);
void rbm_finalize(rbm_ctx *ctx);

/* Message of the last error on this context (ctx == NULL: last error of a failed
 * rbm_init on this thread).  Valid until the next call on the context. */
const char *rbm_last_error(rbm_ctx *ctx);

/* All work of the context is issued on `cuda_stream` (a cudaStream_t; NULL selects
 * the context's own stream).  Lets a host time the library with its own events. */
int rbm_set_stream(rbm_ctx *ctx, void *cuda_stream);
int rbm_synchronize(rbm_ctx *ctx);
int rbm_device_info(rbm_ctx *ctx, int *sm_count, int64_t *free_bytes, int64_t *total_bytes);

/* Device memory for hosts without a CUDA binding of their own (Julia without CUDA.jl). */
int rbm_device_malloc(rbm_ctx *ctx, int64_t bytes, void **out);
int rbm_device_free(rbm_ctx *ctx, void *ptr);
int rbm_memcpy(rbm_ctx *ctx, void *dst, const void *src, int64_t bytes); /* any direction, synchronous */

/*
 * Page-lock a caller-owned host array so that transfers to and from it run at pinned-memory speed and overlap with
 * the time loop (cudaHostRegister).  Julia `Array`s are pageable: the Julia wrapper registers the Snapshots arrays
 * (SS.X, SS.V, SS.A: src/particles/snapshots.jl:16-25) once after allocating them and unregisters them before they
 * are freed.  Every entry point detects registered memory by itself (no flag to pass).  `bytes` must cover the whole
 * array; registering the same range twice is an error (RBM_ERR_CUDA).
 */
int rbm_host_register(rbm_ctx *ctx, void *ptr, int64_t bytes);
int rbm_host_unregister(rbm_ctx *ctx, void *ptr);

/* Number of kernels this context has launched so far (bench.py: "gpu_launches"). */
int64_t rbm_launch_count(rbm_ctx *ctx);

/* Per-kernel device timing for bench.py's roofline: while enabled, every launch of the named
 * kernels ("k_step", "k_step_save", "k_step_dual", "k_deposit", "k_solve", "k_dgemm") is bracketed by a CUDA-event
 * pair on the context's stream.  rbm_profile_query synchronises and returns the number of launches
 * and their summed duration since rbm_profile_enable(ctx, 1). */
int rbm_profile_enable(rbm_ctx *ctx, int on);
int rbm_profile_query(rbm_ctx *ctx, const char *kernel, int64_t *launches, double *total_ms);

/* ------------------------------------------- multi-GPU (one process per GPU) */

/* NCCL is resolved at run time (dlopen of the NCCL already loaded in the process,
 * else libnccl.so.2).  The id is 128 bytes: rank 0 creates it, the host broadcasts
 * it (torch.distributed / MPI / a file) and every rank calls rbm_comm_init.
 * With a communicator attached, rbm_pic_* treats its particles as this rank's chunk
 * of a global particle set: the charge density and the diagnostics K, M are summed
 * over ranks every step; rbm_gram / rbm_pod_evd sum the Gram matrix over ranks
 * (rows = particle chunks) and rbm_deim reduces its arg-max over ranks. */
int rbm_comm_unique_id(void *id128);
int rbm_comm_init(rbm_ctx *ctx, int nranks, int rank, const void *id128);
int rbm_comm_destroy(rbm_ctx *ctx);

/* Optional fast path for the per-step exchange (the only latency-bound collective of the path): each rank
 * exposes a small buffer through CUDA IPC (64-byte handle), the host all-gathers the handles, and from then on
 * the step kernel of every rank reads the peers' charge-density sums directly over NVLink (P2P loads, sequence
 * flags, fixed summation order) instead of going through ncclAllReduce between two kernels.  If peer access is
 * not possible rbm_comm_peer_attach fails with RBM_ERR_COMM and the NCCL path stays in use.
 * Failure mode: the in-kernel wait for a peer's vector is bounded in wall-clock time (RBM_XCH_TIMEOUT_MS, default
 * 10000).  If a rank never publishes (it died, or entered rbm_pic_run more than that much later), the waiting ranks
 * stop waiting, finish the time loop with incomplete charge sums and rbm_pic_run returns RBM_ERR_COMM: every output
 * of such a run -- including snapshots already copied to host arrays -- must be discarded. */
int rbm_comm_peer_handle(rbm_ctx *ctx, void *handle64);
int rbm_comm_peer_attach(rbm_ctx *ctx, const void *handles /* nranks x 64 bytes, rank order; NULL: switch off */);

/* ------------------------------------------------------------ full-order PIC */

/*
 * Replaces  PoissonSolverPBSplines(p, nh, L)  (scripts/bump_on_tail_1_training.jl:73;
 * fields p, L: src/particles/poisson.jl:7-8) together with the particle storage of
 * ParticleList (test/trainingset_tests.jl:17-21) and the VPIntegratorCache work
 * vectors (scripts/bump_on_tail_1_training.jl:70).
 *   np     particles held by THIS context (its chunk when a communicator is attached)
 *   nh     number of periodic B-spline basis functions / cells, 4 <= nh <= 4096
 *   degree spline degree 1, 2 or 3 (3 is the value every reference script uses and the one the tuned kernels serve;
 *          1 and 2 run on the generic kernels)
 *   L      domain length (2*pi/kappa, scripts/bump_on_tail_1_training.jl:64)
 */
int rbm_pic_create(rbm_ctx *ctx, int64_t np, int nh, int degree, double L, rbm_pic **out);
void rbm_pic_destroy(rbm_pic *pic);

/* Initial condition = `particles` of integrate_vp! (x, v, w of length np; ParticleList
 * fields .x .v .w, src/particles/time_marching.jl:122-124).  w == NULL selects the
 * uniform weight w_uniform (= L/np_global, notebooks/VP-ROM-Projections.ipynb cell 2).
 * The stored initial condition is never modified by rbm_pic_run (the reference re-uses
 * `particles` for every sample: scripts/bump_on_tail_1_training.jl:87-97). */
int rbm_pic_set_particles(rbm_pic *pic, const double *x, const double *v, const double *w,
                          double w_uniform);

/*
 * Replaces the whole sample loop of scripts/bump_on_tail_1_training.jl:87-109:
 *     for p in eachindex(pspace)
 *         efield = ScaledPoissonField(poisson, chi[p])                    (:94)
 *         integrate_vp!(particles, efield, lparams, IP, IC; save=true)    (:97)
 *         SS.X[1,:,:,p] .= IC.X ; V ; A ; Phi ; SS.W[:,p] .= IC.W ; K ; M  (:100-108)
 * with all `nparam` samples advanced in lock-step inside one kernel per time step.
 * Time loop = src/particles/time_marching.jl:119-149 with Psi = I:
 *     dt_eff = dt*chi; x += 0.5*dt_eff*v; a = E(x); v += dt_eff*a; x += 0.5*dt_eff*v;
 *     save at step 0 and whenever it % (nt div (ns-1)) == 0.
 * Outputs (any may be NULL = not wanted), layout of Snapshots
 * (src/particles/snapshots.jl:16-25), particle index fastest:
 *     X, V, A : np * ns * nparam      X[i + np*(ts + ns*p)]
 *     Phi     : nh * ns * nparam
 *     W, K, M : ns * nparam
 * Requires ns >= 2 and nt >= ns-1 (the reference divides by ns-1).
 * Completion: if any output is a host array the call returns when all outputs are complete.  If every output is
 * a device pointer (or NULL) the call returns once the work is queued on the context's stream; results are then
 * stream-ordered like any other device work (use rbm_synchronize, or consume them on the same stream).
 */
int rbm_pic_run(rbm_pic *pic, const double *chi, int nparam, double dt, int nt, int ns,
                double *X, double *V, double *A, double *Phi, double *W, double *K, double *M);

/*
 * Output conventions that nothing in the reference checkout pins (the arithmetic lives in PoissonSolvers.jl /
 * VlasovMethods.jl), as switches, so that the result of tools/reference_vectors.jl +
 * tests/test_reference_vectors.py can be applied without a code change:
 *   phi_sign   +1: the stored coefficients (Phi of rbm_pic_run, rbm_field_coefficients) are those of phi with
 *              a = +d(phi)/dx (default); -1: those of -phi
 *   tap_shift  cyclic shift of the stored coefficient vector (default 0: tap 0 belongs to the particle's own cell)
 *   gauge      constant added to the zero-mean coefficients (default 0)
 *   a_at_save_post_update   0 (default): A at a save holds the acceleration used in that step's kick;
 *              1: the field of the update! at the saved positions, gathered there (one extra pass over x per save)
 * X, V, W, K, M and the dynamics do not depend on any of them.
 */
int rbm_pic_set_conventions(rbm_pic *pic, double phi_sign, int tap_shift, double gauge, int a_at_save_post_update);

/*
 * ElectricField protocol of VlasovMethods as extended in
 * src/particles/electric_field.jl:22-35 and used at src/particles/time_marching.jl:95-99,140:
 *     update!(f, x, w, t)   -> rbm_field_update   (deposit + Poisson solve, phi = phi0/chi^2)
 *     efield!(f, E, x)      -> rbm_field_eval     (E_p = d(phi)/dx at x_p)
 *     energy(f)             -> rbm_field_energy   (chi^2 * 1/2 phi' K phi)
 *     coefficients(f)       -> rbm_field_coefficients (nh spline coefficients)
 * A Julia `B200PoissonField <: ElectricField` forwards to these four, so
 * DEIMElectricField(B200PoissonField(...), ...) works unchanged (electric_field.jl:39-43).
 * n may differ from the np of rbm_pic_create (DEIM evaluates at k_e points).
 */
int rbm_field_update(rbm_pic *pic, const double *x, const double *w, double w_uniform, int64_t n,
                     double chi);
int rbm_field_eval(rbm_pic *pic, const double *x, int64_t n, double *E);
int rbm_field_energy(rbm_pic *pic, double *W);
int rbm_field_coefficients(rbm_pic *pic, double *phi);

/* Building blocks exposed for the parity tests (same kernels the run uses):
 *   rbm_locate   cell index c = floor(mod(x,L)/h) (bit-exact bar) and offset xi
 *   rbm_deposit  rhs_j = sum_p w_p B_j(x_p)       (PoissonSolvers rhs_particles_PBSBasis,
 *                                                  imported at src/particles/time_marching.jl:3)
 *   rbm_pic_step one drift-kick-drift step in place on caller arrays; returns the step's
 *                rhs (nh), phi (nh) and a (n); any of the three may be NULL. */
int rbm_locate(rbm_pic *pic, const double *x, int64_t n, int32_t *cell, double *xi);
int rbm_deposit(rbm_pic *pic, const double *x, const double *w, double w_uniform, int64_t n,
                double *rhs);
int rbm_pic_step(rbm_pic *pic, double *x, double *v, const double *w, double w_uniform, int64_t n,
                 double chi, double dt, double *rhs, double *phi, double *a);

/* ----------------------------------------- reduced online model with DEIM (next row) */

/*
 * DEIMElectricField(field, Psi_x, Psi_e, Pi_e, X) + reduced_integrate_vp
 * (src/particles/electric_field.jl:2-43, src/particles/time_marching.jl:66-150; driver
 * scripts/bump_on_tail_3_testing.jl:108-138), all samples advanced in lock-step:
 *     z_x = Psi'x0, z_v = Psi'v0;  per step  z_x += dt_eff/2 z_v;
 *     update!: x = Psi z_x (N x k_p GEMM over all samples) -> full deposit -> Poisson solve;
 *     efield!: y = (Pi'Psi) z_x -> e = phi'(y) at the k_e DEIM points -> z_a = (Psi'Psi_e (Pi'Psi_e)^-1) e;
 *     z_v += dt_eff z_a;  z_x += dt_eff/2 z_v;  save_solution: X = Psi z_x, V = Psi z_v, Phi, W, K, M.
 * The particles (x0, v0, w) are those of rbm_pic_set_particles on `pic`.  deim_idx holds the k_e selected
 * rows (0-based) as returned by rbm_deim.  Outputs (any may be NULL), column-major:
 *     Zx, Zv : k_p * ns * nparam      X, V : N * ns * nparam      Phi : nh * ns * nparam      W, K, M : ns * nparam
 * (the reference's save_solution does not write SS.A either).
 */
typedef struct rbm_reduced rbm_reduced;
int rbm_reduced_create(rbm_pic *pic, const double *Psi_p, int64_t ldp, int kp, const double *Psi_e, int64_t lde,
                       int ke, const int64_t *deim_idx, rbm_reduced **out);
void rbm_reduced_destroy(rbm_reduced *rm);
int rbm_reduced_run(rbm_reduced *rm, const double *chi, int nparam, double dt, int nt, int ns, double *Zx,
                    double *Zv, double *X, double *V, double *Phi, double *W, double *K, double *M);

/* ------------------------------------------------------------- POD and DEIM */

/*
 * G = [B_1 B_2 ... B_nb]' [B_1 ... B_nb]   (C x C, C = sum ncols, column-major, full
 * symmetric matrix).  Replaces `XV' * XV` of the cotangent lift
 * (src/algorithms/evd.jl:32-33; two blocks X, V, never concatenated) and `S' * S`
 * (:65; one block).  Block b is nrows x ncols[b], column-major with leading
 * dimension ld[b] >= nrows.  With a communicator attached the result is summed over
 * ranks (rows = particle chunks).
 */
int rbm_gram(rbm_ctx *ctx, const double *const *blocks, const int64_t *ncols, const int64_t *ld,
             int nblocks, int64_t nrows, double *G);

/*
 * get_PODBasis_cotangentLiftEVD(X, V; k, tolerance)  (src/algorithms/evd.jl:26-55; two
 * blocks) and get_PODBasis_EVD(S; k, tolerance) (:63-87; one block):
 *   G = S'S; eigen(G); sort by |lambda| descending (src/eigen.jl:4-7);
 *   k == 0: first k with sum_{i<=k} Lambda_i / sum(Lambda) >= 1 - tolerance (:38-46);
 *   Psi = S * Omega[:,1:k] * diag(|Lambda[1:k]|^-1/2)   (:49, :81).
 * Lambda receives ALL C eigenvalues (:54).  *k_out receives the rank.  Psi
 * (nrows x k, leading dimension ldpsi) may be NULL to query k only (nothing is kept: a second
 * call repeats the Gram and the eigensolve -- use rbm_pod_factor / rbm_pod_basis to avoid that); if non-NULL and
 * k == 0 it must have room for psi_capacity columns (RBM_ERR_INVALID if k_out exceeds it).
 */
int rbm_pod_evd(rbm_ctx *ctx, const double *const *blocks, const int64_t *ncols, const int64_t *ld,
                int nblocks, int64_t nrows, int k, double tolerance, double *Lambda, int *k_out,
                double *Psi, int64_t ldpsi, int psi_capacity);

/*
 * The same computation in two explicit steps, for callers that must know the rank before they can allocate Psi
 * (a Julia `Matrix{Float64}(undef, N, k)`): rbm_pod_factor does Gram + eigen + sorteigen + rank and returns an
 * opaque decomposition handle; rbm_pod_basis forms Psi = S Omega[:,1:k] |Lambda|^-1/2 from it (k == 0: the rank
 * rbm_pod_factor chose); rbm_evd_destroy frees it.  The handle owns the C x C eigenvectors in HBM and nothing else:
 * it holds no reference to the caller's blocks, which are passed again to rbm_pod_basis and must still hold the
 * data that was factored (the library cannot check that).  rbm_pod_evd itself keeps no state between calls.
 * Errors: RBM_ERR_NUMERIC when sum(Lambda) is not positive and finite (all-zero snapshots; the reference would
 * produce NaN columns) or when one of the k leading eigenvalues is zero.
 */
typedef struct rbm_evd rbm_evd;
int rbm_pod_factor(rbm_ctx *ctx, const double *const *blocks, const int64_t *ncols, const int64_t *ld,
                   int nblocks, int64_t nrows, int k, double tolerance, double *Lambda, int *k_out,
                   rbm_evd **out);
int rbm_pod_basis(rbm_evd *f, const double *const *blocks, const int64_t *ncols, const int64_t *ld,
                  int nblocks, int64_t nrows, int k, double *Psi, int64_t ldpsi);
void rbm_evd_destroy(rbm_evd *f);

/*
 * get_DEIM_interpolation_matrix(Psi)  (src/algorithms/deim.jl:6-21): greedy DEIM with
 * the reference's SIGNED arg-max (:8, :16).  Returns the k selected rows (0-based,
 * global row numbers when a communicator is attached) instead of the dense N x k
 * 0/1 matrix; Pi[idx[i], i] = 1 rebuilds it.
 */
int rbm_deim(rbm_ctx *ctx, const double *Psi, int64_t n, int64_t ld, int k, int64_t *idx);

/*
 * Snapshot projection  Z = Psi' * S  (k x m), as in
 * scripts/bump_on_tail_2_projections.jl:44-60 (X~ = Psi_p' X, A~ = Psi_p' A) and the
 * reduced initial condition of src/particles/time_marching.jl:122-123.
 */
int rbm_project(rbm_ctx *ctx, const double *Psi, int64_t n, int64_t ldpsi, int k, const double *S,
                int64_t lds, int64_t m, double *Z, int64_t ldz);

/* ------------------------------ sampler and growth-rate regression on the device (next row 4) */

/*
 * Initial-condition draw of the bump-on-tail distribution
 *     f(x, v) = (1 + eps cos(kappa x)) [(1-a) N(0,1)(v) + a N(v0, sigma^2)(v)],  x in [0, 2 pi / kappa)
 * replacing VlasovMethods.BumpOnTail.draw_accept_reject(np, params) as called at
 * scripts/bump_on_tail_1_training.jl:76-79 (Random.seed!(1234) + accept-reject on the host).
 * Counter-based: particle p = first + i draws from Philox4x32-10 with key = seed and counter
 * (p, draw number), so the result does not depend on the launch geometry or on how the particles
 * are split over ranks (`first` = global index of this rank's first particle).  NOT the Julia
 * MersenneTwister stream.  x, v (and wout, may be NULL; filled with w) are host or device arrays of
 * n doubles; *rejected (may be NULL) returns the number of rejected candidates.
 */
int rbm_sample_bump_on_tail(rbm_ctx *ctx, int64_t n, int64_t first, uint64_t seed, double kappa, double eps,
                            double a, double v0, double sigma, double w, double *x, double *v, double *wout,
                            int64_t *rejected);

/*
 * get_regression_alpha_beta(t, W, n, inds) with find_maxima(w, n)  (src/regression.jl:4-24):
 * for every column of W (ns x nparam, column-major, leading dimension ldw) the least-squares
 * line alpha + beta t through log W at the local maxima selected by inds (0-based positions in
 * the list of maxima; a maximum needs n non-decreasing samples before and n non-increasing after).
 * nmaxima (may be NULL) returns how many maxima each trace has, up to the largest one asked for.
 * RBM_ERR_INVALID where the reference throws BoundsError (fewer maxima than inds selects).
 */
int rbm_growth_rate(rbm_ctx *ctx, const double *t, const double *W, int64_t ns, int64_t nparam, int64_t ldw,
                    int n, const int *inds, int ninds, double *alpha, double *beta, int *nmaxima);

#ifdef __cplusplus
}
#endif
#endif /* RBM_B200_H */
