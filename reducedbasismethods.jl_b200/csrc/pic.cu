// pic.cu -- host side of the full-order PIC path behind the C ABI (include/rbm_b200.h):
// launch geometry, sample grouping for HBM capacity, the time loop of
// src/particles/time_marching.jl:119-149 (reference), snapshot hand-over in the layout of
// src/particles/snapshots.jl:16-25, and the ElectricField protocol of
// src/particles/electric_field.jl:22-35.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "pic_kernels.cuh"

using namespace rbm;

struct rbm_pic {
    rbm_ctx *ctx = nullptr;
    int64_t np = 0;
    Geom g{};
    // initial condition (device copies)
    DevBuf x0, v0, w;
    bool has_w = false, has_particles = false;
    double w_uniform = 0.0;
    DevBuf pinv;  // [nh]
    // launch geometry of the histogram kernels
    int T = 0;           // threads per block
    int rep = 1;         // gather-table replication
    bool ghist = false;  // large-nh fallback
    size_t smem_step = 0, smem_dep = 0;
    // saved steps: k_step<DUAL> carries a second histogram for the update! at the saved positions.  dual = 1: two
    // thread-private histograms (both fit: small nh); 2: two histograms with paired columns (lanes l, l+16 share one:
    // the shared memory of ONE thread-private histogram); 0: separate deposit pass
    int dual = 0;
    size_t smem_dual = 0;
    // main (unsaved) step kernel: normally (T, rep) with a thread-private histogram; Tm != T: more threads with paired
    // histogram columns (HM = 3)
    int Tm = 0, repm = 0, hm_main = 0;
    size_t smem_main = 0;
    // workspaces of rbm_pic_run, kept between calls (grown on demand, never shrunk)
    DevBuf ws_mid2;  // second deposit-partial buffer (ping-pong between consecutive steps)
    DevBuf ws_acc;   // [3][group][nh] fixed-point deposit sums (one rank, uniform weights, nh <= 64), rotating over the steps
    DevBuf ws_tick;  // [group + 1] arrival counters of the fused publish (zero between launches)
    DevBuf ws_tick2; // [nsamp] arrival counters of the deposit kernel's fused field solve (zero between launches)
    DevBuf ws_red;   // [group][nh] deposit sums over chunks and ranks (multi-rank runs)
    DevBuf ws_x, ws_v, ws_par, ws_mid, ws_save, ws_km, ws_coef, ws_coef2, ws_scr, ws_scr2, ws_gX, ws_gV, ws_gA, ws_Phi, ws_WKM;
    // CUDA graph of the time loop: a repeated rbm_pic_run with identical shapes and pointers replays ONE graph
    // launch instead of ~nt kernel launches, which makes the run immune to host-side launch jitter.
    struct GraphKey {
        const void *X, *V, *A, *x, *v, *w, *Phi;
        int64_t np;
        int nparam, nt, ns, nchunks, has_w, ws_gen, p0, ng, nranks, peer;   // 10 ints: no padding, so memcmp is exact
        double dt, w_uniform;
        bool operator==(const GraphKey &o) const { return memcmp(this, &o, sizeof(GraphKey)) == 0; }
    };
    // one slot per lock-step sample group of a run (a run of 64 samples in groups of 4 replays 16 graphs)
    struct GraphSlot {
        GraphKey key{}, seen{};
        bool seen_valid = false;
        cudaGraphExec_t exec = nullptr;
        int64_t launches = 0;
    };
    std::vector<GraphSlot> graphs;
    int ws_gen = 0;  // bumped whenever a workspace is reallocated (invalidates the graphs)
    int64_t mem_key[6] = {0, 0, 0, 0, 0, 0};   // shape for which mem_budget was determined
    size_t mem_budget = 0;
    // parameters (chi, dt) currently held in ws_par: an identical call skips the upload and its host synchronisation
    std::vector<double> par_chi;
    double par_dt = 0.0;
    int par_gen = -1;
    // staged (host) snapshots leave the device on a second stream, slot by slot, while the time loop runs
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> slot_ev;
    // two sets of staging buffers alternate between consecutive lock-step groups (when HBM allows): set_done[s] marks
    // the end of the copies out of set s, which the group after next waits for before it writes the set again
    cudaEvent_t set_done[2] = {nullptr, nullptr};
    ~rbm_pic()
    {
        for (auto &g : graphs)
            if (g.exec) cudaGraphExecDestroy(g.exec);
        for (auto e : slot_ev) cudaEventDestroy(e);
        for (auto e : set_done)
            if (e) cudaEventDestroy(e);
        if (copy_stream) cudaStreamDestroy(copy_stream);
    }
    // Output conventions that nothing in the reference pins (rbm_pic_set_conventions).  Defaults: the
    // stored coefficients are those of phi with a = +phi', zero mean, tap 0 = the particle's own cell, and A at a save is
    // the acceleration used in that step's kick.
    double conv_phi_sign = 1.0, conv_gauge = 0.0;
    int conv_tap_shift = 0;
    bool conv_a_post = false;
    DevBuf ws_conv;  // scratch of the Phi output transform
    // single-sample field state (ElectricField protocol)
    DevBuf f_phi, f_coef, f_scal;  // phi[nh], coef[nh*3], scal = {chi, W}
    // persistent scratch of the protocol calls (update!/efield! run once per time step in a reduced_integrate_vp
    // driven from Julia: no cudaMalloc/cudaFree per call)
    DevBuf f_part, f_scr, f_in_x, f_in_w, f_out;
    bool f_valid = false;
    double f_chi = 1.0, f_W = 0.0;
};

namespace {

bool host_is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

bool getenv_on(const char *name, bool dflt)
{
    const char *v = getenv(name);
    if (!v || !*v) return dflt;
    return v[0] != '0';
}

template <class K> int set_smem(rbm_ctx *ctx, K kernel, size_t bytes)
{
    RBM_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return RBM_OK;
}

// first row of the circulant pseudo-inverse of the periodic B-spline stiffness matrix (zero-mean gauge), via its Fourier
// symbol: degree 3: K_s = (1/h) circ(-1,-24,-15,80,-15,-24,-1)/120; degree 2: (1/h) circ(-1,-2,6,-2,-1)/6; degree 1:
// (1/h) circ(-1,2,-1).
void stiffness_pinv_row(int nh, double h, int degree, std::vector<double> &row)
{
    row.assign(nh, 0.0);
    const long double PI = 3.14159265358979323846264338327950288L;
    std::vector<long double> lam(nh, 0.0L);
    for (int k = 1; k < nh; ++k) {
        long double th = 2.0L * PI * (long double)k / (long double)nh;
        if (degree == 3)
            lam[k] = (80.0L - 30.0L * cosl(th) - 48.0L * cosl(2.0L * th) - 2.0L * cosl(3.0L * th)) / (120.0L * (long double)h);
        else if (degree == 2)
            lam[k] = (6.0L - 4.0L * cosl(th) - 2.0L * cosl(2.0L * th)) / (6.0L * (long double)h);
        else
            lam[k] = (2.0L - 2.0L * cosl(th)) / (long double)h;
    }
    for (int j = 0; j < nh; ++j) {
        long double acc = 0.0L;
        for (int k = 1; k < nh; ++k) {
            // reduce j*k mod nh before the cosine to keep the argument small
            long double th = 2.0L * PI * (long double)(((int64_t)j * k) % nh) / (long double)nh;
            acc += cosl(th) / lam[k];
        }
        row[j] = (double)(acc / (long double)nh);
    }
}

// a slice of a persistent workspace
struct View {
    void *p;
    size_t bytes;
    template <class T> T *as() const { return static_cast<T *>(p); }
};

struct Chunks {
    int nchunks = 1;
    int64_t chunk_len = 0;
    int grid = 1;
};

// Partition np particles of each of nsamp samples into work items (sample, chunk).  Chunks are a
// whole number of 4T-particle tiles (the vector path of k_step), so only the last chunk of a sample
// has a scalar remainder; with nSM chunks every block gets the same number of tiles (+-1).
Chunks plan_chunks(const rbm_pic *pic, int64_t np, int nsamp)
{
    Chunks c;
    const int nsm = pic->ctx->sm_count;
    const int64_t tile = 4 * (int64_t)std::max(pic->T, pic->Tm);
    const int64_t ntiles = std::max<int64_t>(1, (np + tile - 1) / tile);
    // one work item per SM when possible: every item pays a fixed cost (field solve, histogram publish),
    // so a sample is cut into nSM/nsamp chunks, not more
    int64_t want = nsamp >= nsm ? 1 : nsm / std::max(nsamp, 1);
    const int nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, want));
    const int64_t tiles_per_chunk = (ntiles + nchunks - 1) / nchunks;
    c.chunk_len = tiles_per_chunk * tile;
    c.nchunks = (int)((np + c.chunk_len - 1) / c.chunk_len);
    if (c.nchunks < 1) c.nchunks = 1;
    c.grid = std::min(nsm, nsamp * c.nchunks);
    return c;
}

// How many samples advance in lock-step, given the cap from HBM capacity and item length.  Every launch runs
// `g x chunks(g)` work items on nSM blocks; a group size that leaves SMs without an item (7 samples x 19 chunks = 133 of
// 148) or a ragged last group costs more than it looks, so the size is chosen by a small cost model over g <= cap:
// launches x (longest block's particles + a fixed per-item cost of ~10k particle-equivalents: histogram zeroing, in-kernel
// field solve, publish).
int plan_group(const rbm_pic *pic, int64_t np, int cap, int nparam)
{
    const int nsm = pic->ctx->sm_count;
    const int64_t tile = 4 * (int64_t)std::max(pic->T, pic->Tm);
    const int64_t ntiles = std::max<int64_t>(1, (np + tile - 1) / tile);
    cap = std::max(1, std::min(cap, nparam));
    const double fixed = 10000.0;
    auto launch_cost = [&](int g) -> double {      // one launch of g samples
        if (g <= 0) return 0.0;
        const int64_t chunks = std::max<int64_t>(1, std::min<int64_t>(ntiles, g >= nsm ? 1 : nsm / g));
        const int64_t tiles_per_chunk = (ntiles + chunks - 1) / chunks;
        const int64_t nchunks = (ntiles + tiles_per_chunk - 1) / tiles_per_chunk;
        const int64_t items = (int64_t)g * nchunks;
        const int64_t rounds = (items + nsm - 1) / nsm;
        return (double)rounds * ((double)(tiles_per_chunk * tile) + fixed);
    };
    int best = 1;
    double best_cost = 1e300;
    for (int g = 1; g <= cap; ++g) {
        const double cost = (double)(nparam / g) * launch_cost(g) + launch_cost(nparam % g);
        if (cost < best_cost * (1.0 - 1e-9)) { best_cost = cost; best = g; }
    }
    return best;
}

// (T, REP) instantiations of the shared-memory-histogram kernels
#define RBM_FOR_CONFIG(T_, REP_, BODY)                          \
    if (pic->T == T_ && pic->rep == REP_) {                     \
        constexpr int TT = T_, RR = REP_;                       \
        (void)RR;                                               \
        BODY                                                    \
    }

template <class K> int launch_step_kernel(rbm_pic *pic, K kern, const StepArgs &a, int grid, int threads, size_t smem)
{
    rbm_ctx *ctx = pic->ctx;
    // programmatic dependent launch: consecutive steps overlap launch latency and prologue
    // (not while profiling: the event pairs serialise the stream anyway)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    static const bool pdl_on = !(getenv("RBM_PDL") && getenv("RBM_PDL")[0] == '0');
    cfg.numAttrs = (ctx->profiling || !pdl_on) ? 0 : 1;
    RBM_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, a));
    return RBM_OK;
}

template <int TT, int RR> int step_variant(rbm_pic *pic, const StepArgs &a, int grid, bool save, bool wt, bool configure)
{
    rbm_ctx *ctx = pic->ctx;
    auto go = [&](auto kern) -> int {
        if (configure) return set_smem(ctx, kern, pic->smem_step);
        return launch_step_kernel(pic, kern, a, grid, TT, pic->smem_step);
    };
    if (configure) {
        RBM_TRY(go(k_step<TT, 4, RR, false, false, 0>));
        RBM_TRY(go(k_step<TT, 4, RR, true, false, 0>));
        RBM_TRY(go(k_step<TT, 4, RR, false, true, 0>));
        return go(k_step<TT, 4, RR, true, true, 0>);
    }
    if (save) return wt ? go(k_step<TT, 4, RR, true, true, 0>) : go(k_step<TT, 4, RR, true, false, 0>);
    return wt ? go(k_step<TT, 4, RR, false, true, 0>) : go(k_step<TT, 4, RR, false, false, 0>);
}

// saved step with the second histogram
template <int TT, int RR, int MODE> int dual_variant(rbm_pic *pic, const StepArgs &a, int grid, bool wt, bool configure)
{
    rbm_ctx *ctx = pic->ctx;
    auto go = [&](auto kern) -> int {
        if (configure) return set_smem(ctx, kern, pic->smem_dual);
        return launch_step_kernel(pic, kern, a, grid, TT, pic->smem_dual);
    };
    if (configure) {
        RBM_TRY(go(k_step<TT, 4, RR, true, false, MODE>));
        return go(k_step<TT, 4, RR, true, true, MODE>);
    }
    return wt ? go(k_step<TT, 4, RR, true, true, MODE>) : go(k_step<TT, 4, RR, true, false, MODE>);
}

template <int TT> int deposit_variant(rbm_pic *pic, const DepositArgs &a, int grid, bool wt, bool configure)
{
    rbm_ctx *ctx = pic->ctx;
    auto go = [&](auto kern) -> int {
        if (configure) return set_smem(ctx, kern, pic->smem_dep);
        kern<<<grid, TT, pic->smem_dep, ctx->stream>>>(a);
        return RBM_OK;
    };
    if (configure) {
        RBM_TRY(go(k_deposit<TT, false>));
        return go(k_deposit<TT, true>);
    }
    return wt ? go(k_deposit<TT, true>) : go(k_deposit<TT, false>);
}

// main step with paired histogram columns (HM = 3)
template <int TT, int RR> int step_variant_paired(rbm_pic *pic, const StepArgs &a, int grid, bool save, bool wt, bool configure)
{
    rbm_ctx *ctx = pic->ctx;
    auto go = [&](auto kern) -> int {
        if (configure) return set_smem(ctx, kern, pic->smem_main);
        return launch_step_kernel(pic, kern, a, grid, TT, pic->smem_main);
    };
    if (configure) {
        RBM_TRY(go(k_step<TT, 4, RR, false, false, 3>));
        RBM_TRY(go(k_step<TT, 4, RR, true, false, 3>));
        RBM_TRY(go(k_step<TT, 4, RR, false, true, 3>));
        return go(k_step<TT, 4, RR, true, true, 3>);
    }
    if (save) return wt ? go(k_step<TT, 4, RR, true, true, 3>) : go(k_step<TT, 4, RR, true, false, 3>);
    return wt ? go(k_step<TT, 4, RR, false, true, 3>) : go(k_step<TT, 4, RR, false, false, 3>);
}

int dispatch_step(rbm_pic *pic, const StepArgs &a, int grid, bool save, bool wt, bool configure)
{
    if (pic->hm_main == 3) {
        if (pic->Tm == 512 && pic->repm == 16) return step_variant_paired<512, 16>(pic, a, grid, save, wt, configure);
        if (pic->Tm == 384 && pic->repm == 8) return step_variant_paired<384, 8>(pic, a, grid, save, wt, configure);
        if (pic->Tm == 192 && pic->repm == 4) return step_variant_paired<192, 4>(pic, a, grid, save, wt, configure);
        return rbm_fail(pic->ctx, RBM_ERR_UNSUPPORTED, "no paired-column instantiation for T = %d", pic->Tm);
    }
    RBM_FOR_CONFIG(512, 16, return (step_variant<TT, RR>(pic, a, grid, save, wt, configure));)
    RBM_FOR_CONFIG(384, 16, return (step_variant<TT, RR>(pic, a, grid, save, wt, configure));)
    RBM_FOR_CONFIG(192, 8, return (step_variant<TT, RR>(pic, a, grid, save, wt, configure));)
    RBM_FOR_CONFIG(96, 4, return (step_variant<TT, RR>(pic, a, grid, save, wt, configure));)
    return rbm_fail(pic->ctx, RBM_ERR_UNSUPPORTED, "no kernel instantiation for T = %d", pic->T);
}

int dispatch_dual(rbm_pic *pic, const StepArgs &a, int grid, bool wt, bool configure)
{
    if (pic->dual == 1) { RBM_FOR_CONFIG(512, 16, return (dual_variant<TT, RR, 1>(pic, a, grid, wt, configure));) }
    if (pic->dual == 2) {
        RBM_FOR_CONFIG(512, 16, return (dual_variant<TT, RR, 2>(pic, a, grid, wt, configure));)
        RBM_FOR_CONFIG(384, 16, return (dual_variant<TT, RR, 2>(pic, a, grid, wt, configure));)
        RBM_FOR_CONFIG(192, 8, return (dual_variant<TT, RR, 2>(pic, a, grid, wt, configure));)
        RBM_FOR_CONFIG(96, 4, return (dual_variant<TT, RR, 2>(pic, a, grid, wt, configure));)
    }
    return rbm_fail(pic->ctx, RBM_ERR_UNSUPPORTED, "no dual-histogram instantiation for T = %d (mode %d)", pic->T, pic->dual);
}

int dispatch_deposit(rbm_pic *pic, const DepositArgs &a, int grid, bool wt, bool configure)
{
    RBM_FOR_CONFIG(512, 16, return (deposit_variant<TT>(pic, a, grid, wt, configure));)
    RBM_FOR_CONFIG(384, 16, return (deposit_variant<TT>(pic, a, grid, wt, configure));)
    RBM_FOR_CONFIG(192, 8, return (deposit_variant<TT>(pic, a, grid, wt, configure));)
    RBM_FOR_CONFIG(96, 4, return (deposit_variant<TT>(pic, a, grid, wt, configure));)
    return rbm_fail(pic->ctx, RBM_ERR_UNSUPPORTED, "no kernel instantiation for T = %d", pic->T);
}

int configure_kernels(rbm_pic *pic)
{
    StepArgs sa{};
    DepositArgs da{};
    RBM_TRY(dispatch_step(pic, sa, 0, false, false, true));
    if (pic->dual) RBM_TRY(dispatch_dual(pic, sa, 0, false, true));
    return dispatch_deposit(pic, da, 0, false, true);
}

}  // namespace

extern "C" int rbm_pic_create(rbm_ctx *ctx, int64_t np, int nh, int degree, double L, rbm_pic **out)
{
    if (!ctx || !out) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pic_create: NULL argument");
    *out = nullptr;
    if (np < 1) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pic_create: np = %lld < 1", (long long)np);
    if (nh < 4 || nh > 4096)
        return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pic_create: nh = %d outside [4, 4096]", nh);
    if (!(L > 0.0) || !std::isfinite(L))
        return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pic_create: domain length L must be positive");
    if (degree < 1 || degree > 3)
        return rbm_fail(ctx, RBM_ERR_UNSUPPORTED,
                        "rbm_pic_create: spline degree %d not implemented (1, 2 and 3 are; every reference script uses 3)", degree);
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    rbm_pic *pic = new (std::nothrow) rbm_pic();
    if (!pic) return rbm_fail(ctx, RBM_ERR_NOMEM, "rbm_pic_create: out of host memory");
    pic->ctx = ctx;
    pic->np = np;
    Geom &g = pic->g;
    g.L = L;
    g.nh = nh;
    g.h = L / (double)nh;
    g.invL = 1.0 / L;
    g.invh = 1.0 / g.h;
    g.degree = degree;
    {
        // The fast locate is certified for |x| < 2^50 L (magic-number rounding range).  Markstein's
        // division correction needs RN(1/h); its one exception is a divisor whose mantissa is all
        // ones -- then xhi_max = 0 sends every particle through the exact path.
        uint64_t bits;
        memcpy(&bits, &g.h, sizeof(bits));
        const bool all_ones = (bits & 0xFFFFFFFFFFFFFull) == 0xFFFFFFFFFFFFFull;
        const double xmax = std::ldexp(L, 50);
        uint64_t xb;
        memcpy(&xb, &xmax, sizeof(xb));
        g.xhi_max = (all_ones || !std::isfinite(xmax)) ? 0 : (int)(xb >> 32);
    }
    // launch geometry: thread-private histogram hist[nh+3][T] + replicated gather table.
    // (T, REP) is a compile-time pair so that all shared-memory offsets are immediates.
    const int avail = ctx->smem_optin;
    auto step_smem = [&](int T, int rep) { return ((size_t)(nh + 3) * T + (size_t)nh * 3 * rep + 64) * 8; };
    auto fits = [&](int T, int rep) { return step_smem(T, rep) <= (size_t)avail; };
    int T = 0, rep = 1;
    if (fits(512, 16)) { T = 512; rep = 16; }
    else if (fits(384, 16)) { T = 384; rep = 16; }
    else if (fits(192, 8)) { T = 192; rep = 8; }
    else if (fits(96, 4)) { T = 96; rep = 4; }
    if (degree != 3) T = 0;   // the shared-memory kernels are cubic only: degrees 1 and 2 take the generic kernels
    if (T == 0) {  // histogram does not fit (or degree != 3): generic kernels with global RED.ADD
        pic->ghist = true;
        T = 256;
        pic->smem_step = 0;
        pic->smem_dep = 0;
    } else {
        pic->smem_step = step_smem(T, rep);
        pic->smem_dep = ((size_t)(nh + 3) * T + 64) * 8;
        // second histogram for saved steps: two thread-private ones if both fit (T = 512, small nh), else paired columns
        const size_t two_full = (2 * (size_t)(nh + 3) * T + (size_t)nh * 3 * rep + 64) * 8;
        if (T == 512 && two_full <= (size_t)avail) { pic->dual = 1; pic->smem_dual = two_full; }
        else { pic->dual = 2; pic->smem_dual = pic->smem_step; }   // 2 x (nh+3) x T/2 columns: the same bytes
        if (!getenv_on("RBM_DUAL_SAVE", true)) pic->dual = 0;
    }
    pic->T = T;
    pic->rep = rep;
    pic->Tm = T; pic->repm = rep; pic->hm_main = 0; pic->smem_main = pic->smem_step;
    if (!pic->ghist) {
        // Paired histogram columns (HM = 3) halve the histogram's shared memory, i.e. twice the threads for a given nh.
        // Measured at np = 1e7: nh = 64, 384 -> 512 threads: 59.7 -> 63.4 us per step (the extra predicated LDS/STS and the
        // two __syncwarp cost more than 16 instead of 12 warps bring; off unless RBM_MAIN_PAIRED=1); large grids, where the
        // thread-private layout leaves only 6 or 3 warps per SM, are the other way round (RBM_MAIN_PAIRED=0 switches it off).
        auto need = [&](int Tm, int repm) { return ((size_t)(nh + 3) * (Tm / 2) + (size_t)nh * 3 * repm + 64) * 8; };
        int Tm = 0, repm = 0;
        if (T == 384 && rep == 16 && getenv_on("RBM_MAIN_PAIRED", false)) { Tm = 512; repm = 16; }
        else if (T == 192 && rep == 8 && getenv_on("RBM_MAIN_PAIRED", true)) { Tm = 384; repm = 8; }
        else if (T == 96 && rep == 4 && getenv_on("RBM_MAIN_PAIRED", true)) { Tm = 192; repm = 4; }
        if (Tm && need(Tm, repm) <= (size_t)avail) { pic->Tm = Tm; pic->repm = repm; pic->hm_main = 3; pic->smem_main = need(Tm, repm); }
    }
    int rc = RBM_OK;
    do {
        if (!pic->ghist && (rc = configure_kernels(pic))) break;
        if ((rc = set_smem(ctx, k_solve, sizeof(double) * solve_smem_doubles(nh, 256)))) break;
        if ((rc = set_smem(ctx, k_solve_slots, sizeof(double) * solve_smem_doubles(nh, 256)))) break;
        std::vector<double> row;
        stiffness_pinv_row(nh, g.h, degree, row);
        cudaError_t e;
        if ((e = pic->pinv.alloc(sizeof(double) * nh)) != cudaSuccess ||
            (e = pic->f_phi.alloc(sizeof(double) * nh)) != cudaSuccess ||
            (e = pic->f_coef.alloc(sizeof(double) * nh * 3)) != cudaSuccess ||
            (e = pic->f_scal.alloc(sizeof(double) * 4)) != cudaSuccess ||
            (e = cudaMemcpy(pic->pinv.p, row.data(), sizeof(double) * nh, cudaMemcpyHostToDevice)) != cudaSuccess) {
            rc = rbm_fail(ctx, RBM_ERR_CUDA, "rbm_pic_create: %s", cudaGetErrorString(e));
            break;
        }
    } while (0);
    if (rc) {
        delete pic;
        return rc;
    }
    *out = pic;
    return RBM_OK;
}

extern "C" void rbm_pic_destroy(rbm_pic *pic)
{
    if (!pic) return;
    cudaSetDevice(pic->ctx->device);
    cudaStreamSynchronize(pic->ctx->stream);
    delete pic;
}

extern "C" int rbm_pic_set_particles(rbm_pic *pic, const double *x, const double *v, const double *w,
                                     double w_uniform)
{
    if (!pic) return RBM_ERR_INVALID;
    rbm_ctx *ctx = pic->ctx;
    if (!x || !v) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pic_set_particles: x and v are required");
    if (!w && !(w_uniform > 0.0))
        return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pic_set_particles: w == NULL needs w_uniform > 0");
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = sizeof(double) * (size_t)pic->np;
    if (!pic->x0.p) RBM_CUDA(ctx, pic->x0.alloc(bytes));
    if (!pic->v0.p) RBM_CUDA(ctx, pic->v0.alloc(bytes));
    RBM_CUDA(ctx, cudaMemcpyAsync(pic->x0.p, x, bytes, cudaMemcpyDefault, ctx->stream));
    RBM_CUDA(ctx, cudaMemcpyAsync(pic->v0.p, v, bytes, cudaMemcpyDefault, ctx->stream));
    if (w) {
        if (!pic->w.p) RBM_CUDA(ctx, pic->w.alloc(bytes));
        RBM_CUDA(ctx, cudaMemcpyAsync(pic->w.p, w, bytes, cudaMemcpyDefault, ctx->stream));
        pic->has_w = true;
        pic->w_uniform = 0.0;
    } else {
        pic->w.release();
        pic->has_w = false;
        pic->w_uniform = w_uniform;
    }
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // caller's host arrays may go away
    pic->has_particles = true;
    return RBM_OK;
}

namespace {
// out[r][j] = sign * in[r][(j - shift) mod nh] + gauge   (np.roll(phi, shift) of the oracle's Conventions.stored_phi)
__global__ void k_phi_convention(const double *in, double *out, int64_t rows, int nh, double sign, int shift, double gauge)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * nh) return;
    const int64_t r = idx / nh;
    const int j = (int)(idx - r * nh);
    int src = (j - shift) % nh;
    if (src < 0) src += nh;
    out[idx] = sign * in[r * nh + src] + gauge;
}
}  // namespace

extern "C" int rbm_pic_set_conventions(rbm_pic *pic, double phi_sign, int tap_shift, double gauge, int a_at_save_post_update)
{
    if (!pic) return RBM_ERR_INVALID;
    if (!(phi_sign == 1.0 || phi_sign == -1.0) || !std::isfinite(gauge))
        return rbm_fail(pic->ctx, RBM_ERR_INVALID, "rbm_pic_set_conventions: phi_sign must be +1 or -1, gauge finite");
    pic->conv_phi_sign = phi_sign;
    pic->conv_tap_shift = tap_shift;
    pic->conv_gauge = gauge;
    pic->conv_a_post = a_at_save_post_update != 0;
    pic->ws_gen++;   // a captured time loop was recorded under the old conventions
    return RBM_OK;
}

namespace {

// deposit of `x` (+ hdt*v) for nsamp samples into `part`; part_km optional
int launch_deposit(rbm_pic *pic, const double *x, const double *v, const double *w, const double *hdt,
                   int64_t np, int64_t ldx, int nsamp, const Chunks &ch, double *part, double *part_km,
                   const SolveArgs *fused_solve = nullptr)
{
    rbm_ctx *ctx = pic->ctx;
    DepositArgs a{};
    a.x = x; a.v = v; a.w = w; a.hdt = hdt; a.part = part; a.part_km = part_km;
    a.np = np; a.ldx = ldx; a.chunk_len = ch.chunk_len; a.nsamp = nsamp; a.nchunks = ch.nchunks;
    a.g = pic->g;
    auto al16 = [](const void *q) { return ((uintptr_t)q & 15) == 0; };
    a.vec = al16(x) && al16(v) && al16(w) && (nsamp == 1 || ldx % 2 == 0) && ch.chunk_len % 4 == 0;
    if (fused_solve) {   // caller guarantees: one rank, shared-memory histogram, ws_tick2 holds nsamp zeroed counters
        a.fuse_solve = 1;
        a.sv = *fused_solve;
        a.tickets = pic->ws_tick2.as<int>();
    }
    if (pic->ghist) {
        RBM_CUDA(ctx, cudaMemsetAsync(part, 0, sizeof(double) * (size_t)nsamp * pic->g.nh, ctx->stream));
        if (part_km)
            RBM_CUDA(ctx, cudaMemsetAsync(part_km, 0, sizeof(double) * (size_t)nsamp * ch.nchunks * 2, ctx->stream));
        int grid = (int)std::min<int64_t>((np + 255) / 256, (int64_t)ctx->sm_count * 8);
        if (w) k_deposit_global<true><<<grid, 256, 0, ctx->stream>>>(a);
        else k_deposit_global<false><<<grid, 256, 0, ctx->stream>>>(a);
    } else {
        ProfScope ps(ctx, "k_deposit");
        dispatch_deposit(pic, a, ch.grid, w != nullptr, false);
    }
    RBM_LAUNCH_CHECK(ctx);
    return RBM_OK;
}

int launch_step(rbm_pic *pic, const StepArgs &a, bool save, const Chunks &ch, bool dual = false)
{
    rbm_ctx *ctx = pic->ctx;
    const bool wt = a.w != nullptr;
    if (dual) {
        ProfScope ps(ctx, "k_step_dual");
        RBM_TRY(dispatch_dual(pic, a, ch.grid, wt, false));
        RBM_LAUNCH_CHECK(ctx);
        return RBM_OK;
    }
    if (pic->ghist) {
        if (a.do_mid)
            RBM_CUDA(ctx, cudaMemsetAsync(a.part, 0, sizeof(double) * (size_t)a.nsamp * pic->g.nh, ctx->stream));
        if (save) {
            if (wt) k_step_global<true, true><<<ch.grid, 256, 0, ctx->stream>>>(a);
            else k_step_global<true, false><<<ch.grid, 256, 0, ctx->stream>>>(a);
        } else {
            if (wt) k_step_global<false, true><<<ch.grid, 256, 0, ctx->stream>>>(a);
            else k_step_global<false, false><<<ch.grid, 256, 0, ctx->stream>>>(a);
        }
    } else {
        ProfScope ps(ctx, save ? "k_step_save" : "k_step");
        RBM_TRY(dispatch_step(pic, a, ch.grid, save, wt, false));
    }
    RBM_LAUNCH_CHECK(ctx);
    return RBM_OK;
}

// Poisson solve for nsamp samples from deposit partials.  With a communicator the
// partials (and K/M partials) are first summed over chunks and then over ranks.
struct SolveOut {
    double *coef = nullptr, *rhs = nullptr, *phi = nullptr, *W = nullptr, *K = nullptr, *M = nullptr;
    int64_t phi_stride = 0, w_stride = 0;
};

// weights of the deposit being solved: an array was used (weighted) or the uniform value wu
struct Weights {
    bool weighted;
    double wu;
};
Weights weights_of(const rbm_pic *pic) { return Weights{pic->has_w, pic->w_uniform}; }

// bound of the in-kernel wait for a peer's charge-density vector (wall clock; RBM_XCH_TIMEOUT_MS, default 10 s)
unsigned long long xch_timeout_ns()
{
    static const unsigned long long ns = [] {
        const char *e = getenv("RBM_XCH_TIMEOUT_MS");
        const double ms = e && *e ? atof(e) : 10000.0;
        return (unsigned long long)(std::max(ms, 1.0) * 1e6);
    }();
    return ns;
}

SolveArgs fill_solve_args(const rbm_pic *pic, const double *part, int part_chunks, const double *part_km, int km_chunks,
                          const double *chi, const SolveOut &o, Weights wt)
{
    SolveArgs a{};
    a.part = part; a.part_km = part_km; a.nchunks = part_chunks; a.km_chunks = km_chunks;
    a.scale = wt.weighted ? 1.0 / 6.0 : wt.wu / 6.0;
    a.km_scale = wt.weighted ? 1.0 : wt.wu;
    a.pinv = pic->pinv.as<double>();
    a.chi = chi;
    a.coef = o.coef; a.rhs_out = o.rhs; a.phi_out = o.phi; a.phi_stride = o.phi_stride;
    a.W_out = o.W; a.K_out = o.K; a.M_out = o.M; a.w_stride = o.w_stride;
    a.nh = pic->g.nh; a.h = pic->g.h;
    a.xch_timeout_ns = xch_timeout_ns();
    a.degree = pic->g.degree;
    return a;
}

int launch_solve(rbm_pic *pic, const double *part, int part_chunks, const double *part_km, int km_chunks,
                 const double *chi, int nsamp, const SolveOut &o, double *comm_scratch, Weights wt)
{
    rbm_ctx *ctx = pic->ctx;
    const int nh = pic->g.nh;
    if (ctx->nranks > 1) {
        // scratch layout: [nsamp][nh] rhs sums | [nsamp][2] K/M sums
        double *r = comm_scratch, *km = comm_scratch + (size_t)nsamp * nh;
        int n = nsamp * nh;
        k_reduce_part<<<(n + 255) / 256, 256, 0, ctx->stream>>>(part, nsamp, part_chunks, nh, r);
        RBM_LAUNCH_CHECK(ctx);
        size_t count = (size_t)n;
        if (part_km) {
            k_reduce_part<<<(nsamp * 2 + 255) / 256, 256, 0, ctx->stream>>>(part_km, nsamp, km_chunks, 2, km);
            RBM_LAUNCH_CHECK(ctx);
            count += (size_t)nsamp * 2;
        }
        RBM_TRY(rbm_allreduce_sum_f64(ctx, r, count));
        part = r; part_chunks = 1;
        part_km = part_km ? km : nullptr; km_chunks = 1;
    }
    const SolveArgs a = fill_solve_args(pic, part, part_chunks, part_km, km_chunks, chi, o, wt);
    // 256 threads = G groups x nh bins (G*nh <= 256 doubles of group partials)
    ProfScope ps(ctx, "k_solve");
    k_solve<<<nsamp, 256, sizeof(double) * solve_smem_doubles(nh, 256), ctx->stream>>>(a);
    RBM_LAUNCH_CHECK(ctx);
    return RBM_OK;
}

// update!(field, x): deposit and field solve.  On one rank with the shared-memory histogram the solve runs in the tail
// of the deposit kernel (the CTA that publishes a sample's last chunk); otherwise deposit, [rank reduction], k_solve.
// part_km_dep: K/M partials written by THIS deposit (needs v) or NULL; part_km_solve/km_chunks: the partials the solve
// reduces into K, M (those of this deposit, or the ones a preceding k_step<SAVE> left).
int launch_deposit_solve(rbm_pic *pic, const double *x, const double *v, const double *w, const double *hdt, int64_t np,
                         int64_t ldx, int nsamp, const Chunks &ch, double *part, double *part_km_dep,
                         const double *part_km_solve, int km_chunks, const double *chi, const SolveOut &o,
                         double *comm_scratch, Weights wt)
{
    rbm_ctx *ctx = pic->ctx;
    const bool fused = ctx->nranks == 1 && !pic->ghist && getenv_on("RBM_FUSE_SAVE_SOLVE", true);
    if (!fused) {
        RBM_TRY(launch_deposit(pic, x, v, w, hdt, np, ldx, nsamp, ch, part, part_km_dep));
        return launch_solve(pic, part, pic->ghist ? 1 : ch.nchunks, part_km_solve, km_chunks, chi, nsamp, o, comm_scratch, wt);
    }
    if (pic->ws_tick2.bytes < sizeof(int) * (size_t)nsamp) {
        if (pic->ws_tick2.p) RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        pic->ws_gen++;   // a captured time loop holds the old pointer
        RBM_CUDA(ctx, pic->ws_tick2.alloc(sizeof(int) * (size_t)nsamp));
        RBM_CUDA(ctx, cudaMemsetAsync(pic->ws_tick2.p, 0, pic->ws_tick2.bytes, ctx->stream));
    }
    const SolveArgs sv = fill_solve_args(pic, part, ch.nchunks, part_km_solve, km_chunks, chi, o, wt);
    return launch_deposit(pic, x, v, w, hdt, np, ldx, nsamp, ch, part, part_km_dep, &sv);
}

}  // namespace

extern "C" int rbm_pic_run(rbm_pic *pic, const double *chi, int nparam, double dt, int nt, int ns,
                           double *X, double *V, double *A, double *Phi, double *W, double *K, double *M)
{
    if (!pic) return RBM_ERR_INVALID;
    rbm_ctx *ctx = pic->ctx;
    if (!pic->has_particles) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pic_run: call rbm_pic_set_particles first");
    if (!chi || nparam < 1) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pic_run: need nparam >= 1 values of chi");
    if (ns < 2 || nt < ns - 1)
        return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pic_run: need ns >= 2 and nt >= ns-1 (nt = %d, ns = %d)", nt, ns);
    if (!(dt > 0.0)) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pic_run: dt must be positive");
    for (int p = 0; p < nparam; ++p)
        if (!(chi[p] > 0.0) || !std::isfinite(chi[p]))
            return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pic_run: chi[%d] must be positive and finite", p);
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));

    const int64_t np = pic->np;
    const int nh = pic->g.nh;
    const int nsave = nt / (ns - 1);  // src/particles/time_marching.jl:119
    const int64_t ldp = (np + 3) & ~(int64_t)3;
    const double *w = pic->has_w ? pic->w.as<double>() : nullptr;

    // ---- how many samples advance in lock-step: bounded by HBM (state + staged snapshots)
    const bool stageX = X && !rbm_is_device_ptr(X), stageV = V && !rbm_is_device_ptr(V),
               stageA = A && !rbm_is_device_ptr(A);
    const size_t snap_bytes_per_sample = sizeof(double) * (size_t)np * ns;
    const bool staged = stageX || stageV || stageA;
    if (staged) {
        if (!pic->copy_stream) RBM_CUDA(ctx, cudaStreamCreateWithFlags(&pic->copy_stream, cudaStreamNonBlocking));
        while ((int)pic->slot_ev.size() < ns) {
            cudaEvent_t e;
            RBM_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            pic->slot_ev.push_back(e);
        }
        for (auto &e : pic->set_done)
            if (!e) RBM_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    // no return path (errors included) leaves copies to the caller's host arrays in flight
    struct CopyGuard {
        cudaStream_t s;
        ~CopyGuard() { if (s) cudaStreamSynchronize(s); }
    } copy_guard{staged ? pic->copy_stream : nullptr};
    size_t per_sample = sizeof(double) * 2 * (size_t)ldp +
                        ((stageX ? 1 : 0) + (stageV ? 1 : 0) + (stageA ? 1 : 0)) * snap_bytes_per_sample;
    // HBM budget of the lock-step group.  cudaMemGetInfo is a slow driver call (and what this handle already holds
    // must count as available, or an identical second call would plan a smaller group): asked once per shape.
    const int64_t mem_key[6] = {np, nparam, ns, stageX, stageV, stageA};
    if (memcmp(mem_key, pic->mem_key, sizeof(mem_key)) != 0 || pic->mem_budget == 0) {
        size_t free_b = 0, total_b = 0;
        RBM_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
        free_b += pic->ws_x.bytes + pic->ws_v.bytes + pic->ws_gX.bytes + pic->ws_gV.bytes + pic->ws_gA.bytes;
        pic->mem_budget = (size_t)(0.85 * (double)free_b);
        memcpy(pic->mem_key, mem_key, sizeof(mem_key));
    }
    const size_t budget = pic->mem_budget;
    int group = (int)std::min<size_t>((size_t)nparam, budget / std::max<size_t>(per_sample, 1));
    {
        // How many samples advance in lock-step.  One work item (a chunk of one sample) per SM and launch; every item
        // pays ~9 us of fixed cost (histogram zeroing, in-kernel field solve, publish), so items should be long, but
        // launches whose state is many times the L2 lose the alternating-sweep reuse and measured slower again.
        // Measured optimum (np = 1e6 / 2.5e6 / 1e7 per sample: 16 / 16 / 2 samples per launch): about 128 Ki particles
        // per item -- 6.0-6.2 us per 1e6 particle-steps, against 7.3-8.1 with the earlier "state fits L2" rule.
        const double target_item = 131072.0;
        int by_item = (int)std::max(1.0, std::floor((double)ctx->sm_count * target_item / (double)std::max<int64_t>(np, 1) + 0.5));
        int cap = std::max(1, std::min(group, by_item));
        if (const char *e = getenv("RBM_GROUP")) {   // experiments: force the lock-step group size
            const int gforce = atoi(e);
            if (gforce > 0) cap = std::max(1, std::min(group, gforce));
        }
        group = plan_group(pic, np, cap, nparam);
    }

    Chunks ch = plan_chunks(pic, np, group);
    int max_items = group * ch.nchunks;
    if (nparam % group) {   // the last, smaller group is cut into more chunks per sample
        const int last = nparam % group;
        max_items = std::max(max_items, last * plan_chunks(pic, np, last).nchunks);
    }
    const size_t part_items = pic->ghist ? (size_t)group : (size_t)max_items;  // (sample, chunk) partial slots

    // persistent workspaces (cudaMalloc of the 160 MB state would otherwise cost as much as several steps)
    auto grow = [&](DevBuf &b, size_t bytes) -> cudaError_t {
        if (b.bytes >= bytes && b.p) return cudaSuccess;
        cudaStreamSynchronize(ctx->stream);
        pic->ws_gen++;
        return b.alloc(bytes);
    };
    RBM_CUDA(ctx, grow(pic->ws_mid2, sizeof(double) * part_items * nh));
    RBM_CUDA(ctx, grow(pic->ws_acc, sizeof(unsigned long long) * 3 * (size_t)group * nh));
    if (pic->ws_tick.bytes < sizeof(int) * ((size_t)group + 1)) {
        RBM_CUDA(ctx, grow(pic->ws_tick, sizeof(int) * ((size_t)group + 1)));
        RBM_CUDA(ctx, cudaMemsetAsync(pic->ws_tick.p, 0, pic->ws_tick.bytes, ctx->stream));
    }
    if (pic->ws_tick2.bytes < sizeof(int) * (size_t)group) {
        RBM_CUDA(ctx, grow(pic->ws_tick2, sizeof(int) * (size_t)group));
        RBM_CUDA(ctx, cudaMemsetAsync(pic->ws_tick2.p, 0, pic->ws_tick2.bytes, ctx->stream));
    }
    RBM_CUDA(ctx, grow(pic->ws_red, sizeof(double) * (size_t)group * nh));
    RBM_CUDA(ctx, grow(pic->ws_x, sizeof(double) * (size_t)group * ldp));
    RBM_CUDA(ctx, grow(pic->ws_v, sizeof(double) * (size_t)group * ldp));
    RBM_CUDA(ctx, grow(pic->ws_par, sizeof(double) * 4 * (size_t)nparam));
    RBM_CUDA(ctx, grow(pic->ws_mid, sizeof(double) * part_items * nh));
    // saved steps keep their deposit and K/M partials (one slot per saved step) and are solved together after the loop
    const bool defer_saves = pic->dual != 0 && !pic->ghist && !pic->conv_a_post && getenv_on("RBM_DEFER_SAVE_SOLVE", true);
    const size_t save_slots = defer_saves ? (size_t)ns : 1;
    RBM_CUDA(ctx, grow(pic->ws_save, sizeof(double) * save_slots * part_items * nh));
    RBM_CUDA(ctx, grow(pic->ws_km, sizeof(double) * save_slots * (size_t)max_items * 2));
    if (defer_saves && ctx->nranks > 1) RBM_CUDA(ctx, grow(pic->ws_scr2, sizeof(double) * (size_t)ns * group * (nh + 2)));
    RBM_CUDA(ctx, grow(pic->ws_coef, sizeof(double) * (size_t)group * nh * 3));
    RBM_CUDA(ctx, grow(pic->ws_coef2, sizeof(double) * (size_t)group * nh * 3));
    RBM_CUDA(ctx, grow(pic->ws_scr, sizeof(double) * (size_t)group * (nh + 2)));
    // Staging buffers of the host snapshots.  With more than one group and room in HBM there are two sets, used in
    // turn: the copies of a group's last slots (which nothing in that group's time loop can hide) then run under the
    // time loop of the next group instead of holding it up.
    const size_t stage_set = snap_bytes_per_sample * group;   // bytes of one set, per array
    const int nsets = (staged && nparam > group && getenv_on("RBM_STAGE_SETS2", true) &&
                       (size_t)group * (2 * per_sample - sizeof(double) * 2 * (size_t)ldp) <= budget) ? 2 : 1;
    if (stageX) RBM_CUDA(ctx, grow(pic->ws_gX, stage_set * nsets));
    if (stageV) RBM_CUDA(ctx, grow(pic->ws_gV, stage_set * nsets));
    if (stageA) RBM_CUDA(ctx, grow(pic->ws_gA, stage_set * nsets));
    RBM_CUDA(ctx, grow(pic->ws_Phi, sizeof(double) * (size_t)nparam * ns * nh));
    RBM_CUDA(ctx, grow(pic->ws_WKM, sizeof(double) * 3 * (size_t)nparam * ns));
    DevBuf &sx = pic->ws_x, &sv = pic->ws_v, &part_mid = pic->ws_mid, &part_save = pic->ws_save, &part_km = pic->ws_km,
           &coef = pic->ws_coef, &coef_save = pic->ws_coef2, &scratch = pic->ws_scr, &gX = pic->ws_gX, &gV = pic->ws_gV,
           &gA = pic->ws_gA;
    double *par = pic->ws_par.as<double>();
    const View d_chi{par, sizeof(double) * nparam}, d_hdt{par + nparam, sizeof(double) * nparam},
        d_dte{par + 2 * (size_t)nparam, sizeof(double) * nparam}, d_zero{par + 3 * (size_t)nparam, sizeof(double) * nparam};
    const View d_Phi{pic->ws_Phi.p, sizeof(double) * (size_t)nparam * ns * nh};
    double *wkm = pic->ws_WKM.as<double>();
    const View d_W{wkm, sizeof(double) * (size_t)nparam * ns}, d_K{wkm + (size_t)nparam * ns, sizeof(double) * (size_t)nparam * ns},
        d_M{wkm + 2 * (size_t)nparam * ns, sizeof(double) * (size_t)nparam * ns};
    RBM_CUDA(ctx, cudaMemsetAsync(d_Phi.p, 0, d_Phi.bytes, ctx->stream));
    RBM_CUDA(ctx, cudaMemsetAsync(wkm, 0, 3 * d_W.bytes, ctx->stream));
    RBM_CUDA(ctx, cudaMemsetAsync(d_zero.p, 0, d_zero.bytes, ctx->stream));

    std::vector<double> h_hdt(nparam), h_dte(nparam);
    for (int p = 0; p < nparam; ++p) {
        h_dte[p] = dt * chi[p];        // time_marching.jl:127
        h_hdt[p] = 0.5 * h_dte[p];     // `0.5 .* dt_eff .* v` = (0.5*dt_eff)*v  (:137)
    }
    const bool same_par = pic->par_gen == pic->ws_gen && pic->par_dt == dt && (int)pic->par_chi.size() == nparam &&
                          std::equal(pic->par_chi.begin(), pic->par_chi.end(), chi);
    if (!same_par) {
        RBM_CUDA(ctx, cudaMemcpyAsync(d_chi.p, chi, sizeof(double) * nparam, cudaMemcpyHostToDevice, ctx->stream));
        RBM_CUDA(ctx, cudaMemcpyAsync(d_hdt.p, h_hdt.data(), sizeof(double) * nparam, cudaMemcpyHostToDevice, ctx->stream));
        RBM_CUDA(ctx, cudaMemcpyAsync(d_dte.p, h_dte.data(), sizeof(double) * nparam, cudaMemcpyHostToDevice, ctx->stream));
        RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // host vectors above are temporaries
        pic->par_chi.assign(chi, chi + nparam);
        pic->par_dt = dt;
        pic->par_gen = pic->ws_gen;
    }

    if (ctx->nranks > 1 && ctx->xch_ready && getenv_on("RBM_PEER_XCH", true)) {
        // Rank barrier on the device before the first peer wait: a rank that enters rbm_pic_run late (pageable staging
        // copies, first-call module load) is waited for inside NCCL, which has no timeout, instead of inside the bounded
        // spin of the step kernel.
        RBM_CUDA(ctx, cudaMemsetAsync(pic->ws_scr.p, 0, sizeof(double), ctx->stream));
        RBM_TRY(rbm_allreduce_sum_f64(ctx, pic->ws_scr.as<double>(), 1));
    }
    const int64_t snap_stride = (int64_t)np * ns;
    for (int p0 = 0; p0 < nparam; p0 += group) {
        const int ng = std::min(group, nparam - p0);
        if (ng != group) ch = plan_chunks(pic, np, ng);
        const int pchunks = pic->ghist ? 1 : ch.nchunks;
        // where this group's snapshots go (device memory either way)
        const int set = nsets == 2 ? (p0 / group) & 1 : 0;
        const size_t set_off = (size_t)set * stage_set / sizeof(double);
        double *Xg = X ? (stageX ? gX.as<double>() + set_off : X + (size_t)p0 * snap_stride) : nullptr;
        double *Vg = V ? (stageV ? gV.as<double>() + set_off : V + (size_t)p0 * snap_stride) : nullptr;
        double *Ag = A ? (stageA ? gA.as<double>() + set_off : A + (size_t)p0 * snap_stride) : nullptr;
        // the copies out of this set (two groups ago) must have finished before this group writes it
        if (staged && nsets == 2 && p0 >= 2 * group) RBM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, pic->set_done[set], 0));
        const double *chi_g = d_chi.as<double>() + p0, *hdt_g = d_hdt.as<double>() + p0,
                     *dte_g = d_dte.as<double>() + p0;
        double *Phi_g = d_Phi.as<double>() + (size_t)p0 * ns * nh;
        double *W_g = d_W.as<double>() + (size_t)p0 * ns, *K_g = d_K.as<double>() + (size_t)p0 * ns,
               *M_g = d_M.as<double>() + (size_t)p0 * ns;
        auto snap_vec_ok = [&](int ts) {
            auto ok = [&](double *b) {
                return !b || (((uintptr_t)(b + (size_t)np * ts) & 15) == 0 && (snap_stride & 1) == 0);
            };
            return (ok(Xg) && ok(Vg) && ok(Ag)) ? 1 : 0;
        };

        // ---- every sample starts from the same particles (bump_on_tail_1_training.jl:87-97)
        {
            dim3 grid((unsigned)std::min<int64_t>((np + 255) / 256, (int64_t)ctx->sm_count * 8), ng);
            ProfScope ps(ctx, "k_broadcast");
            k_broadcast<<<grid, 256, 0, ctx->stream>>>(pic->x0.as<double>(), pic->v0.as<double>(),
                                                       sx.as<double>(), sv.as<double>(), np, ldp);
            RBM_LAUNCH_CHECK(ctx);
        }
        // ---- initial save (time_marching.jl:130 -> save_solution :81-105)
        int ts = 0;
        {
            SolveOut o;
            o.coef = coef_save.as<double>();
            o.phi = Phi_g; o.phi_stride = (int64_t)ns * nh;
            o.W = W_g; o.K = K_g; o.M = M_g; o.w_stride = ns;
            RBM_TRY(launch_deposit_solve(pic, sx.as<double>(), sv.as<double>(), w, d_zero.as<double>(), np, ldp, ng, ch,
                                         part_save.as<double>(), part_km.as<double>(), part_km.as<double>(), ch.nchunks,
                                         chi_g, o, scratch.as<double>(), weights_of(pic)));
        }
        if (Xg || Vg || Ag) {
            GatherArgs ga{};
            ga.x = sx.as<double>(); ga.v = sv.as<double>(); ga.coef = coef_save.as<double>();
            ga.A = Ag; ga.X = Xg; ga.V = Vg; ga.snap_stride = snap_stride;
            ga.np = np; ga.ldx = ldp; ga.nsamp = ng; ga.g = pic->g;
            dim3 grid((unsigned)std::min<int64_t>((np + 255) / 256, (int64_t)ctx->sm_count * 8), ng);
            {
                ProfScope ps(ctx, "k_gather");
                k_gather<<<grid, 256, 0, ctx->stream>>>(ga);
            }
            RBM_LAUNCH_CHECK(ctx);
            if (staged) RBM_CUDA(ctx, cudaEventRecord(pic->slot_ev[0], ctx->stream));
        }
        ts = 1;
        // ---- deposit of the first half-drifted positions -> field of step 1.
        // The deposit partials ping-pong between two buffers: step `it` solves its field from the
        // partials of the previous launch (cur) and publishes the next ones (nxt).
        double *pbuf[2] = {part_mid.as<double>(), pic->ws_mid2.as<double>()};
        int cur = 0;
        RBM_TRY(launch_deposit(pic, sx.as<double>(), sv.as<double>(), w, hdt_g, np, ldp, ng, ch, pbuf[cur], nullptr));
        // what the next k_step solves from: the chunk partials, or (multi-rank) their sum over chunks and ranks
        const bool fuse = !pic->ghist;
        const double *solve_part = pbuf[cur];
        int solve_chunks = pchunks;
        PeerXch peers{};   // filled by rank_reduce when the peer-memory exchange is in use
        const auto use_peer = [&]() { return ctx->nranks > 1 && ctx->xch_ready && ng * nh <= XCH_CAPR && getenv_on("RBM_PEER_XCH", true); };
        // the exchange-buffer slot the NEXT published vector goes to (k_step publishes from its own tail)
        unsigned long long xch_rel = 0;   // exchanges published so far in this group's time loop
        unsigned long long *const xch_epoch_ptr =
            ctx->xch_local ? (unsigned long long *)((char *)ctx->xch_local + sizeof(double) * 2 * XCH_R * XCH_CAPR) + 2 * XCH_R + 1 : nullptr;
        auto next_pub = [&]() -> PeerXch {
            PeerXch pb{};
            for (int r = 0; r < ctx->nranks; ++r) pb.base[r] = (double *)ctx->xch_peer[r];
            pb.nranks = ctx->nranks; pb.self = ctx->rank; pb.epoch = xch_epoch_ptr; pb.seq = xch_rel + 1;
            return pb;
        };
        // published == true: a k_step launch already wrote + flagged the vector (fused publish)
        auto rank_reduce = [&](const double *part, bool published) -> int {
            peers.nranks = 0;
            if (ctx->nranks == 1) {
                solve_part = part;
                solve_chunks = pchunks;
                return RBM_OK;
            }
            const int n = ng * nh;
            if (use_peer()) {
                // P2P path: this rank's sums go to its exchange buffer; the next k_step reads every rank's buffer over NVLink
                peers = next_pub();
                ++xch_rel;
                if (!published) {
                    k_reduce_publish<<<1, 1024, 0, ctx->stream>>>(part, ng, pchunks, nh, peers);
                    RBM_LAUNCH_CHECK(ctx);
                }
                solve_part = part;   // unused by the peer path
                solve_chunks = pchunks;
                return RBM_OK;
            }
            double *r = pic->ws_red.as<double>();
            k_reduce_part<<<(n + 255) / 256, 256, 0, ctx->stream>>>(part, ng, pchunks, nh, r);
            RBM_LAUNCH_CHECK(ctx);
            RBM_TRY(rbm_allreduce_sum_f64(ctx, r, (size_t)n));
            solve_part = r;
            solve_chunks = 1;
            return RBM_OK;
        };
        if (fuse) {
            RBM_TRY(rank_reduce(pbuf[cur], false));
        } else {
            SolveOut o;
            o.coef = coef.as<double>();
            RBM_TRY(launch_solve(pic, pbuf[cur], pchunks, nullptr, 0, chi_g, ng, o, scratch.as<double>(), weights_of(pic)));
        }
        // Fixed-point deposit sums (SolveArgs::acc): on one rank with uniform weights and nh <= 64 the step kernel's CTAs add
        // their chunk sums to ONE vector per sample with 64-bit integer atomics instead of publishing 148 partial vectors that
        // every CTA of the next launch reads back (11 MB through the L2 per step, 0.9 us at its full bandwidth).  A chunk sum is
        // at most 4 np (6 B_1(0) = 4 per particle), so 2^e with 4 np 2^e < 2^62 scales it to an integer: the rounding is
        // 2^-e absolute per chunk (2^-36 at np = 1e7, against bin values of 1e6) and the integer sum is exact, i.e.
        // independent of the order in which the CTAs arrive -- bit-reproducible like the fixed-order sum it replaces.
        // Three buffers rotate: step it adds to it % 3, reads (it - 1) % 3 and clears (it + 1) % 3.
        const bool fast_solve = getenv_on("RBM_FAST_SOLVE", true);
        const bool use_acc = fuse && fast_solve && ctx->nranks == 1 && !pic->has_w && nh <= 64 && (nh & 1) == 0 &&
                             pic->g.degree == 3 && getenv_on("RBM_ACC_FIXED", true);
        unsigned long long *const acc_base = pic->ws_acc.as<unsigned long long>();
        const size_t acc_len = (size_t)ng * nh;
        int acc_exp = 0;
        std::frexp(4.0 * (double)np + 1.0, &acc_exp);   // 4 np + 1 < 2^acc_exp
        const double acc_scale = std::ldexp(1.0, 62 - acc_exp), acc_inv = std::ldexp(1.0, acc_exp - 62);
        // round-robin tiles: measured +3 % at nh = 16, neutral at nh = 64 on one rank; with a communicator (many samples in
        // lock-step, few chunks per sample) contiguous chunks were 1.4 % faster at 8 ranks, so that stays the default there
        const int interleave = !pic->ghist && getenv_on("RBM_INTERLEAVE", ctx->nranks == 1) ? 1 : 0;
        // ---- time loop (time_marching.jl:132-149), issued directly or captured once into a CUDA graph
        auto time_loop = [&]() -> int {
            int first_deferred = -1;   // first saved slot whose field solve is deferred to the end of the loop
            if (use_acc) RBM_CUDA(ctx, cudaMemsetAsync(acc_base, 0, sizeof(unsigned long long) * 3 * acc_len, ctx->stream));
            for (int it = 1; it <= nt; ++it) {
                const bool save = (it % nsave == 0) && ts < ns;
                const int nxt = cur ^ 1;
                StepArgs a{};
                a.x = sx.as<double>(); a.v = sv.as<double>(); a.w = w; a.coef = coef.as<double>();
                a.part = pbuf[nxt]; a.part_km = part_km.as<double>();
                if (save) {
                    a.X = Xg ? Xg + (size_t)np * ts : nullptr;
                    a.V = Vg ? Vg + (size_t)np * ts : nullptr;
                    a.A = Ag ? Ag + (size_t)np * ts : nullptr;
                    a.snap_vec = snap_vec_ok(ts);
                }
                a.snap_stride = snap_stride;
                a.hdt = hdt_g; a.dte = dte_g;
                a.np = np; a.ldp = ldp; a.chunk_len = ch.chunk_len; a.nsamp = ng; a.nchunks = ch.nchunks;
                a.do_mid = it < nt ? 1 : 0; a.g = pic->g;
                a.interleave = interleave;
                a.reverse = (it & 1) ? 0 : 1;   // alternate the sweep direction: L2 reuse of the state just written
                if (fuse) {
                    // every block solves its sample's Poisson problem itself (no solve kernel, no serial tail)
                    a.fuse_solve = 1;
                    SolveArgs &sa = a.sv;
                    sa.peers = peers;
                    sa.part = solve_part; sa.nchunks = solve_chunks; sa.part_km = nullptr; sa.km_chunks = 0;
                    sa.scale = pic->has_w ? 1.0 / 6.0 : pic->w_uniform / 6.0;
                    sa.km_scale = 0.0;
                    sa.pinv = pic->pinv.as<double>(); sa.chi = chi_g;
                    sa.nh = nh; sa.h = pic->g.h; sa.degree = pic->g.degree;
                    sa.xch_timeout_ns = xch_timeout_ns();
                    sa.no_fast = fast_solve ? 0 : 1;
                    if (use_acc) {
                        if (it >= 2) { sa.acc = acc_base + (size_t)((it - 1) % 3) * acc_len; sa.acc_inv = acc_inv; }
                        if (a.do_mid) { a.acc_out = acc_base + (size_t)(it % 3) * acc_len; a.acc_scale = acc_scale; }
                        a.acc_zero = acc_base + (size_t)((it + 1) % 3) * acc_len;
                    }
                    if (a.do_mid && use_peer()) {   // this launch publishes the vector the next step consumes
                        a.publish = 1;
                        a.tickets = pic->ws_tick.as<int>();
                        a.pub = next_pub();
                    }
                }
                // saved step: the update! at the saved positions rides along in k_step<DUAL> (second histogram).  Its field
                // solve (Phi, W, K, M of the slot: outputs only, the dynamics do not need them) is deferred: the partials of
                // every saved slot are kept and one kernel solves all slots after the loop, instead of a serial tail in the
                // last block of every saved step.  (Not deferred: convention "post-update A", which needs the field at once.)
                const bool dual = save && pic->dual != 0 && !pic->ghist;
                const size_t slot_part = (size_t)ng * ch.nchunks * nh, slot_km = (size_t)ng * ch.nchunks * 2;
                if (dual) {
                    if (defer_saves) {
                        a.part_save = part_save.as<double>() + (size_t)ts * slot_part;
                        a.part_km = part_km.as<double>() + (size_t)ts * slot_km;
                    } else {
                        a.part_save = part_save.as<double>();
                        if (ctx->nranks == 1) {
                            SolveOut o;
                            o.phi = Phi_g + (size_t)ts * nh; o.phi_stride = (int64_t)ns * nh;
                            o.W = W_g + ts; o.K = K_g + ts; o.M = M_g + ts; o.w_stride = ns;
                            if (pic->conv_a_post && Ag) o.coef = coef_save.as<double>();
                            a.save_solve = 1;
                            a.tickets_save = pic->ws_tick2.as<int>();
                            a.sv_save = fill_solve_args(pic, part_save.as<double>(), ch.nchunks, part_km.as<double>(), ch.nchunks,
                                                        chi_g, o, weights_of(pic));
                        }
                    }
                }
                RBM_TRY(launch_step(pic, a, save, ch, dual));
                const bool a_post = save && pic->conv_a_post && Ag != nullptr;
                if (save && staged && !a_post) RBM_CUDA(ctx, cudaEventRecord(pic->slot_ev[ts], ctx->stream));
                if (save) {
                    // update!(efield, x_saved) for Phi and W (time_marching.jl:95-99)
                    SolveOut o;
                    o.phi = Phi_g + (size_t)ts * nh; o.phi_stride = (int64_t)ns * nh;
                    o.W = W_g + ts; o.K = K_g + ts; o.M = M_g + ts; o.w_stride = ns;
                    if (a_post) o.coef = coef_save.as<double>();
                    if (!dual)
                        RBM_TRY(launch_deposit_solve(pic, sx.as<double>(), nullptr, w, nullptr, np, ldp, ng, ch,
                                                     part_save.as<double>(), nullptr, part_km.as<double>(), ch.nchunks,
                                                     chi_g, o, scratch.as<double>(), weights_of(pic)));
                    else if (ctx->nranks > 1 && !defer_saves)   // partials are there already: sum over chunks and ranks, then solve
                        RBM_TRY(launch_solve(pic, part_save.as<double>(), ch.nchunks, part_km.as<double>(), ch.nchunks, chi_g, ng, o,
                                             scratch.as<double>(), weights_of(pic)));
                    if (dual && defer_saves) { if (first_deferred < 0) first_deferred = ts; }
                    if (a_post) {
                        // convention "post_update": IC.A at a save holds the field of this update! gathered at the saved
                        // positions (one extra pass over x), not the acceleration of the step's kick
                        GatherArgs ga{};
                        ga.x = sx.as<double>(); ga.coef = coef_save.as<double>();
                        ga.A = Ag + (size_t)np * ts; ga.snap_stride = snap_stride;
                        ga.np = np; ga.ldx = ldp; ga.nsamp = ng; ga.g = pic->g;
                        dim3 grid((unsigned)std::min<int64_t>((np + 255) / 256, (int64_t)ctx->sm_count * 8), ng);
                        k_gather<<<grid, 256, 0, ctx->stream>>>(ga);
                        RBM_LAUNCH_CHECK(ctx);
                        if (staged) RBM_CUDA(ctx, cudaEventRecord(pic->slot_ev[ts], ctx->stream));
                    }
                    ++ts;
                }
                if (it < nt) {
                    if (fuse) {
                        RBM_TRY(rank_reduce(pbuf[nxt], a.publish != 0));
                    } else {
                        SolveOut o;
                        o.coef = coef.as<double>();
                        RBM_TRY(launch_solve(pic, pbuf[nxt], pchunks, nullptr, 0, chi_g, ng, o, scratch.as<double>(), weights_of(pic)));
                    }
                }
                cur = nxt;
            }
            if (first_deferred >= 0) {
                // the saved slots first_deferred .. ts-1: one block per (slot, sample)
                const int nslots = ts - first_deferred, nq = nslots * ng;
                const size_t slot_part = (size_t)ng * ch.nchunks * nh, slot_km = (size_t)ng * ch.nchunks * 2;
                const double *pp = part_save.as<double>() + (size_t)first_deferred * slot_part;
                const double *pk = part_km.as<double>() + (size_t)first_deferred * slot_km;
                int pc = ch.nchunks, kc = ch.nchunks;
                if (ctx->nranks > 1) {   // sum over chunks, then ONE all-reduce for all slots
                    double *rr = pic->ws_scr2.as<double>(), *km = rr + (size_t)nq * nh;
                    k_reduce_part<<<(nq * nh + 255) / 256, 256, 0, ctx->stream>>>(pp, nq, pc, nh, rr);
                    RBM_LAUNCH_CHECK(ctx);
                    k_reduce_part<<<(nq * 2 + 255) / 256, 256, 0, ctx->stream>>>(pk, nq, kc, 2, km);
                    RBM_LAUNCH_CHECK(ctx);
                    RBM_TRY(rbm_allreduce_sum_f64(ctx, rr, (size_t)nq * (nh + 2)));
                    pp = rr; pk = km; pc = 1; kc = 1;
                }
                SolveOut o;
                o.phi = Phi_g; o.phi_stride = (int64_t)ns * nh;
                o.W = W_g; o.K = K_g; o.M = M_g; o.w_stride = ns;
                const SolveArgs sa = fill_solve_args(pic, pp, pc, pk, kc, chi_g, o, weights_of(pic));
                ProfScope ps(ctx, "k_solve");
                k_solve_slots<<<nq, 256, sizeof(double) * solve_smem_doubles(nh, 256), ctx->stream>>>(sa, ng, first_deferred);
                RBM_LAUNCH_CHECK(ctx);
            }
            if (xch_rel > 0) {   // the next loop's exchanges are numbered from epoch + xch_rel on every rank
                k_xch_advance<<<1, 32, 0, ctx->stream>>>(xch_epoch_ptr, xch_rel);
                RBM_LAUNCH_CHECK(ctx);
            }
            return RBM_OK;
        };
        // Graph policy: device-resident outputs, not profiling.  The first call with a new (shape, pointer) key runs
        // directly; from the second identical call on, each sample group's loop is one graph launch.  With a
        // communicator the captured loop contains the fused peer exchange (sequence numbers relative to the device-side
        // epoch word, so the kernel arguments do not change between replays) or the ncclAllReduce calls.
        bool graphed = false;
        const int gslot = p0 / group;
        if (!ctx->profiling && !pic->ghist && !staged && getenv_on("RBM_GRAPH", true) && gslot < 4096) {
            if ((int)pic->graphs.size() <= gslot) pic->graphs.resize(gslot + 1);
            rbm_pic::GraphSlot &gs = pic->graphs[gslot];
            rbm_pic::GraphKey key;
            memset(&key, 0, sizeof(key));
            key.X = Xg; key.V = Vg; key.A = Ag; key.x = sx.p; key.v = sv.p; key.w = w; key.Phi = Phi_g;
            key.np = np; key.nparam = nparam; key.nt = nt; key.ns = ns; key.nchunks = ch.nchunks;
            key.has_w = pic->has_w ? 1 : 0; key.ws_gen = pic->ws_gen; key.p0 = p0; key.ng = ng;
            key.nranks = ctx->nranks; key.peer = use_peer() ? 1 : 0; key.dt = dt; key.w_uniform = pic->w_uniform;
            if (gs.exec && key == gs.key) {
                graphed = true;
            } else if (gs.seen_valid && key == gs.seen) {
                if (gs.exec) { cudaGraphExecDestroy(gs.exec); gs.exec = nullptr; }
                const int64_t l0 = ctx->launches;
                const int ts0 = ts, cur0 = cur;
                const double *sp0 = solve_part; const int sc0 = solve_chunks;
                const PeerXch peers0 = peers; const unsigned long long rel0 = xch_rel;
                RBM_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
                int rc = time_loop();
                cudaGraph_t graph = nullptr;
                cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
                gs.launches = ctx->launches - l0;
                ctx->launches = l0;
                // the capture only recorded; nothing ran
                ts = ts0; cur = cur0; solve_part = sp0; solve_chunks = sc0; peers = peers0; xch_rel = rel0;
                cudaError_t ei = cudaErrorUnknown;
                if (rc == RBM_OK && e == cudaSuccess && graph && (ei = cudaGraphInstantiate(&gs.exec, graph, 0)) == cudaSuccess) {
                    gs.key = key;
                    graphed = true;
                } else {
                    cudaGetLastError();
                    gs.exec = nullptr;
                }
                if (getenv_on("RBM_DEBUG", false))
                    fprintf(stderr, "[rbm] graph capture slot %d: rc=%d end=%s inst=%s launches=%lld\n", gslot, rc, cudaGetErrorString(e),
                            cudaGetErrorString(ei), (long long)gs.launches);
                if (graph) cudaGraphDestroy(graph);
            }
            gs.seen = key;
            gs.seen_valid = true;
        }
        if (graphed) {
            RBM_CUDA(ctx, cudaGraphLaunch(pic->graphs[gslot].exec, ctx->stream));
            ctx->launches += pic->graphs[gslot].launches;
        } else {
            RBM_TRY(time_loop());
        }
        // ---- hand the group's snapshots to the host arrays: every saved slot is copied on the copy stream as
        // soon as the kernel that wrote it has finished, concurrently with the rest of the time loop (the
        // whole loop is already enqueued, so this also overlaps when the host arrays are pageable)
        if (staged) {
            const int nslots = ts;   // slots actually written
            for (int s = 0; s < nslots; ++s) {
                RBM_CUDA(ctx, cudaStreamWaitEvent(pic->copy_stream, pic->slot_ev[s], 0));
                const size_t pitch = sizeof(double) * (size_t)snap_stride, width = sizeof(double) * (size_t)np;
                const size_t off = (size_t)np * s;
                // pinned destination: one pitched copy per array; pageable destination: one contiguous copy per
                // sample (the driver stages pitch x height bytes of a pitched copy to pageable memory)
                auto slot_copy = [&](double *host, const double *dev) -> int {
                    if (host_is_pinned(host)) {
                        RBM_CUDA(ctx, cudaMemcpy2DAsync(host + (size_t)p0 * snap_stride + off, pitch, dev + off, pitch, width, ng,
                                                        cudaMemcpyDeviceToHost, pic->copy_stream));
                    } else {
                        for (int q = 0; q < ng; ++q)
                            RBM_CUDA(ctx, cudaMemcpyAsync(host + (size_t)(p0 + q) * snap_stride + off, dev + (size_t)q * snap_stride + off,
                                                          width, cudaMemcpyDeviceToHost, pic->copy_stream));
                    }
                    return RBM_OK;
                };
                if (stageX) RBM_TRY(slot_copy(X, Xg));
                if (stageV) RBM_TRY(slot_copy(V, Vg));
                if (stageA) RBM_TRY(slot_copy(A, Ag));
            }
            if (nsets == 2)
                RBM_CUDA(ctx, cudaEventRecord(pic->set_done[set], pic->copy_stream));
            else
                RBM_CUDA(ctx, cudaStreamSynchronize(pic->copy_stream));   // the staging buffers are reused by the next group
        }
    }
    if (staged) RBM_CUDA(ctx, cudaStreamSynchronize(pic->copy_stream));
    if (ctx->xch_ready) {   // a rank that never published its vector was waited for ~3 s and then flagged
        unsigned long long status = 0;
        RBM_CUDA(ctx, cudaMemcpyAsync(&status, (char *)ctx->xch_local + sizeof(double) * 2 * XCH_R * XCH_CAPR + 2 * XCH_R * sizeof(unsigned long long),
                                      sizeof(status), cudaMemcpyDeviceToHost, ctx->stream));
        RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (status != 0) {
            // re-arm the exchange for later runs; the outputs of THIS run are incomplete
            cudaMemsetAsync((char *)ctx->xch_local + sizeof(double) * 2 * XCH_R * XCH_CAPR + 2 * XCH_R * sizeof(unsigned long long), 0,
                            sizeof(unsigned long long), ctx->stream);
            cudaStreamSynchronize(ctx->stream);
            return rbm_fail(ctx, RBM_ERR_COMM, "rbm_pic_run: timed out waiting for a peer's charge-density vector "
                                               "(RBM_XCH_TIMEOUT_MS); the outputs of this run are incomplete");
        }
    }
    const void *phi_src = d_Phi.p;
    if (Phi && (pic->conv_phi_sign != 1.0 || pic->conv_tap_shift != 0 || pic->conv_gauge != 0.0)) {
        if (pic->ws_conv.bytes < d_Phi.bytes) {
            RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            RBM_CUDA(ctx, pic->ws_conv.alloc(d_Phi.bytes));
        }
        const int64_t rows = (int64_t)nparam * ns;
        k_phi_convention<<<(unsigned)((rows * nh + 255) / 256), 256, 0, ctx->stream>>>(d_Phi.as<double>(), pic->ws_conv.as<double>(), rows, nh,
                                                                                     pic->conv_phi_sign, pic->conv_tap_shift, pic->conv_gauge);
        RBM_LAUNCH_CHECK(ctx);
        phi_src = pic->ws_conv.p;
    }
    if (Phi) RBM_CUDA(ctx, cudaMemcpyAsync(Phi, phi_src, d_Phi.bytes, cudaMemcpyDefault, ctx->stream));
    if (W) RBM_CUDA(ctx, cudaMemcpyAsync(W, d_W.p, d_W.bytes, cudaMemcpyDefault, ctx->stream));
    if (K) RBM_CUDA(ctx, cudaMemcpyAsync(K, d_K.p, d_K.bytes, cudaMemcpyDefault, ctx->stream));
    if (M) RBM_CUDA(ctx, cudaMemcpyAsync(M, d_M.p, d_M.bytes, cudaMemcpyDefault, ctx->stream));
    // Host outputs must be complete on return.  With device-resident outputs only, the results are ordered on the
    // context's stream like any other device work (rbm_synchronize waits): back-to-back runs then queue without the GPU
    // ever waiting for the host.
    auto on_host = [](const void *q) { return q && !rbm_is_device_ptr(q); };
    if (staged || on_host(Phi) || on_host(W) || on_host(K) || on_host(M) || ctx->nranks > 1 || ctx->profiling)
        RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RBM_OK;
}

// ---------------------------------------------------------------- field protocol

extern "C" int rbm_field_update(rbm_pic *pic, const double *x, const double *w, double w_uniform, int64_t n,
                                double chi)
{
    if (!pic) return RBM_ERR_INVALID;
    rbm_ctx *ctx = pic->ctx;
    if (!x || n < 1) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_field_update: need n >= 1 positions");
    if (!w && !(w_uniform > 0.0)) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_field_update: w == NULL needs w_uniform > 0");
    if (!(chi > 0.0)) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_field_update: chi must be positive");
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    DevIn dx, dw;
    RBM_TRY(dx.stage(ctx, x, sizeof(double) * (size_t)n, &pic->f_in_x));
    RBM_TRY(dw.stage(ctx, w, sizeof(double) * (size_t)n, &pic->f_in_w));
    Chunks ch = plan_chunks(pic, n, 1);
    const int pchunks = pic->ghist ? 1 : ch.nchunks;
    DevBuf &part = pic->f_part, &scratch = pic->f_scr;
    if (part.bytes < sizeof(double) * (size_t)pchunks * pic->g.nh || scratch.bytes < sizeof(double) * (size_t)(pic->g.nh + 2)) {
        RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        RBM_CUDA(ctx, part.reserve(sizeof(double) * (size_t)ctx->sm_count * pic->g.nh));
        RBM_CUDA(ctx, scratch.reserve(sizeof(double) * (size_t)(pic->g.nh + 2)));
    }
    RBM_TRY(launch_deposit(pic, dx.as<double>(), nullptr, dw.as<double>(), nullptr, n, 0, 1, ch, part.as<double>(), nullptr));
    // the field's weights are those of THIS call, not of rbm_pic_set_particles
    double *scal = pic->f_scal.as<double>();
    cudaError_t e = cudaMemcpyAsync(scal, &chi, sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    int rc = RBM_OK;
    if (e == cudaSuccess) {
        SolveOut o;
        o.coef = pic->f_coef.as<double>();
        o.phi = pic->f_phi.as<double>(); o.phi_stride = pic->g.nh;
        o.W = scal + 1; o.w_stride = 1;
        rc = launch_solve(pic, part.as<double>(), pchunks, nullptr, 0, scal, 1, o, scratch.as<double>(),
                          Weights{w != nullptr, w ? 0.0 : w_uniform});
    }
    if (e != cudaSuccess) return rbm_fail(ctx, RBM_ERR_CUDA, "rbm_field_update: %s", cudaGetErrorString(e));
    RBM_TRY(rc);
    double hs[2];
    RBM_CUDA(ctx, cudaMemcpyAsync(hs, scal, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    pic->f_chi = chi;
    pic->f_W = hs[1];
    pic->f_valid = true;
    return RBM_OK;
}

extern "C" int rbm_field_eval(rbm_pic *pic, const double *x, int64_t n, double *E)
{
    if (!pic) return RBM_ERR_INVALID;
    rbm_ctx *ctx = pic->ctx;
    if (!pic->f_valid) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_field_eval: call rbm_field_update first");
    if (!x || !E || n < 1) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_field_eval: bad argument");
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    DevIn dx;
    DevOut dE;
    RBM_TRY(dx.stage(ctx, x, sizeof(double) * (size_t)n, &pic->f_in_x));
    RBM_TRY(dE.prepare(ctx, E, sizeof(double) * (size_t)n, false, &pic->f_out));
    GatherArgs ga{};
    ga.x = dx.as<double>(); ga.coef = pic->f_coef.as<double>(); ga.A = dE.as<double>();
    ga.np = n; ga.ldx = 0; ga.nsamp = 1; ga.g = pic->g;
    dim3 grid((unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8), 1);
    k_gather<<<grid, 256, 0, ctx->stream>>>(ga);
    RBM_LAUNCH_CHECK(ctx);
    RBM_TRY(dE.flush(ctx));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RBM_OK;
}

extern "C" int rbm_field_energy(rbm_pic *pic, double *W)
{
    if (!pic || !W) return RBM_ERR_INVALID;
    if (!pic->f_valid) return rbm_fail(pic->ctx, RBM_ERR_INVALID, "rbm_field_energy: call rbm_field_update first");
    *W = pic->f_W;
    return RBM_OK;
}

extern "C" int rbm_field_coefficients(rbm_pic *pic, double *phi)
{
    if (!pic || !phi) return RBM_ERR_INVALID;
    rbm_ctx *ctx = pic->ctx;
    if (!pic->f_valid) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_field_coefficients: call rbm_field_update first");
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *src = pic->f_phi.p;
    if (pic->conv_phi_sign != 1.0 || pic->conv_tap_shift != 0 || pic->conv_gauge != 0.0) {
        if (pic->ws_conv.bytes < sizeof(double) * pic->g.nh) {
            RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            RBM_CUDA(ctx, pic->ws_conv.alloc(sizeof(double) * pic->g.nh));
        }
        k_phi_convention<<<(pic->g.nh + 255) / 256, 256, 0, ctx->stream>>>(pic->f_phi.as<double>(), pic->ws_conv.as<double>(), 1, pic->g.nh,
                                                                           pic->conv_phi_sign, pic->conv_tap_shift, pic->conv_gauge);
        RBM_LAUNCH_CHECK(ctx);
        src = pic->ws_conv.p;
    }
    RBM_CUDA(ctx, cudaMemcpyAsync(phi, src, sizeof(double) * pic->g.nh, cudaMemcpyDefault, ctx->stream));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RBM_OK;
}

// ---------------------------------------------------------------- building blocks

extern "C" int rbm_locate(rbm_pic *pic, const double *x, int64_t n, int32_t *cell, double *xi)
{
    if (!pic) return RBM_ERR_INVALID;
    rbm_ctx *ctx = pic->ctx;
    if (!x || n < 0) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_locate: bad argument");
    if (n == 0) return RBM_OK;
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    DevIn dx;
    DevOut dc, dxi;
    RBM_TRY(dx.stage(ctx, x, sizeof(double) * (size_t)n));
    RBM_TRY(dc.prepare(ctx, cell, sizeof(int32_t) * (size_t)n));
    RBM_TRY(dxi.prepare(ctx, xi, sizeof(double) * (size_t)n));
    int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8);
    k_locate<<<grid, 256, 0, ctx->stream>>>(dx.as<double>(), n, pic->g, dc.as<int32_t>(), dxi.as<double>());
    RBM_LAUNCH_CHECK(ctx);
    RBM_TRY(dc.flush(ctx));
    RBM_TRY(dxi.flush(ctx));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RBM_OK;
}

extern "C" int rbm_deposit(rbm_pic *pic, const double *x, const double *w, double w_uniform, int64_t n,
                           double *rhs)
{
    if (!pic) return RBM_ERR_INVALID;
    rbm_ctx *ctx = pic->ctx;
    if (!x || !rhs || n < 0) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_deposit: bad argument");
    if (!w && !(w_uniform > 0.0)) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_deposit: w == NULL needs w_uniform > 0");
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    const int nh = pic->g.nh;
    DevOut dr;
    RBM_TRY(dr.prepare(ctx, rhs, sizeof(double) * nh, true));
    if (n > 0) {
        DevIn dx, dw;
        RBM_TRY(dx.stage(ctx, x, sizeof(double) * (size_t)n));
        RBM_TRY(dw.stage(ctx, w, sizeof(double) * (size_t)n));
        Chunks ch = plan_chunks(pic, n, 1);
        const int pchunks = pic->ghist ? 1 : ch.nchunks;
        DevBuf part, scratch, dchi, dphi;
        RBM_CUDA(ctx, part.alloc(sizeof(double) * (size_t)pchunks * nh));
        RBM_CUDA(ctx, scratch.alloc(sizeof(double) * (size_t)(nh + 2)));
        RBM_CUDA(ctx, dchi.alloc(sizeof(double)));
        const double one = 1.0;
        RBM_CUDA(ctx, cudaMemcpyAsync(dchi.p, &one, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        RBM_TRY(launch_deposit(pic, dx.as<double>(), nullptr, dw.as<double>(), nullptr, n, 0, 1, ch, part.as<double>(), nullptr));
        SolveOut o;
        o.rhs = dr.as<double>();
        int rc = launch_solve(pic, part.as<double>(), pchunks, nullptr, 0, dchi.as<double>(), 1, o, scratch.as<double>(),
                              Weights{w != nullptr, w ? 0.0 : w_uniform});
        RBM_TRY(rc);
        RBM_TRY(dr.flush(ctx));
        RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return RBM_OK;
    }
    RBM_TRY(dr.flush(ctx));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RBM_OK;
}

extern "C" int rbm_pic_step(rbm_pic *pic, double *x, double *v, const double *w, double w_uniform, int64_t n,
                            double chi, double dt, double *rhs, double *phi, double *a)
{
    if (!pic) return RBM_ERR_INVALID;
    rbm_ctx *ctx = pic->ctx;
    if (!x || !v || n < 1) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pic_step: bad argument");
    if (!w && !(w_uniform > 0.0)) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pic_step: w == NULL needs w_uniform > 0");
    if (!(chi > 0.0) || !(dt > 0.0)) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_pic_step: chi and dt must be positive");
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    const int nh = pic->g.nh;
    const int64_t ldp = (n + 3) & ~(int64_t)3;
    DevBuf sx, sv, part, part2, part_km, coef, scratch, scal;
    DevIn dw;
    DevOut drhs, dphi, da;
    RBM_CUDA(ctx, sx.alloc(sizeof(double) * (size_t)ldp));
    RBM_CUDA(ctx, sv.alloc(sizeof(double) * (size_t)ldp));
    RBM_CUDA(ctx, cudaMemcpyAsync(sx.p, x, sizeof(double) * (size_t)n, cudaMemcpyDefault, ctx->stream));
    RBM_CUDA(ctx, cudaMemcpyAsync(sv.p, v, sizeof(double) * (size_t)n, cudaMemcpyDefault, ctx->stream));
    RBM_TRY(dw.stage(ctx, w, sizeof(double) * (size_t)n));
    RBM_TRY(drhs.prepare(ctx, rhs, sizeof(double) * nh));
    RBM_TRY(dphi.prepare(ctx, phi, sizeof(double) * nh));
    RBM_TRY(da.prepare(ctx, a, sizeof(double) * (size_t)n));
    Chunks ch = plan_chunks(pic, n, 1);
    const int pchunks = pic->ghist ? 1 : ch.nchunks;
    RBM_CUDA(ctx, part.alloc(sizeof(double) * (size_t)pchunks * nh));
    RBM_CUDA(ctx, part2.alloc(sizeof(double) * (size_t)pchunks * nh));
    RBM_CUDA(ctx, part_km.alloc(sizeof(double) * (size_t)ch.nchunks * 2));
    RBM_CUDA(ctx, coef.alloc(sizeof(double) * (size_t)nh * 3));
    RBM_CUDA(ctx, scratch.alloc(sizeof(double) * (size_t)(nh + 2)));
    RBM_CUDA(ctx, scal.alloc(sizeof(double) * 4));
    const double hs[3] = {chi, 0.5 * (dt * chi), dt * chi};
    RBM_CUDA(ctx, cudaMemcpyAsync(scal.p, hs, sizeof(hs), cudaMemcpyHostToDevice, ctx->stream));
    double *d_chi = scal.as<double>(), *d_hdt = d_chi + 1, *d_dte = d_chi + 2;
    int rc = launch_deposit(pic, sx.as<double>(), sv.as<double>(), dw.as<double>(), d_hdt, n, ldp, 1, ch,
                            part.as<double>(), nullptr);
    if (rc == RBM_OK) {
        SolveOut o;
        o.coef = coef.as<double>(); o.rhs = drhs.as<double>(); o.phi = dphi.as<double>(); o.phi_stride = nh;
        rc = launch_solve(pic, part.as<double>(), pchunks, nullptr, 0, d_chi, 1, o, scratch.as<double>(),
                          Weights{w != nullptr, w ? 0.0 : w_uniform});
    }
    if (rc == RBM_OK) {
        StepArgs sa{};
        sa.x = sx.as<double>(); sa.v = sv.as<double>(); sa.w = dw.as<double>(); sa.coef = coef.as<double>();
        sa.part = part2.as<double>(); sa.part_km = part_km.as<double>();
        sa.A = da.as<double>(); sa.snap_stride = 0; sa.snap_vec = 0;
        sa.hdt = d_hdt; sa.dte = d_dte;
        sa.np = n; sa.ldp = ldp; sa.chunk_len = ch.chunk_len; sa.nsamp = 1; sa.nchunks = ch.nchunks;
        sa.do_mid = 1; sa.g = pic->g;
        rc = launch_step(pic, sa, true, ch);
    }
    RBM_TRY(rc);
    RBM_CUDA(ctx, cudaMemcpyAsync(x, sx.p, sizeof(double) * (size_t)n, cudaMemcpyDefault, ctx->stream));
    RBM_CUDA(ctx, cudaMemcpyAsync(v, sv.p, sizeof(double) * (size_t)n, cudaMemcpyDefault, ctx->stream));
    RBM_TRY(drhs.flush(ctx));
    RBM_TRY(dphi.flush(ctx));
    RBM_TRY(da.flush(ctx));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RBM_OK;
}

// ============================================================================================
// Reduced online model with DEIM (reference: src/particles/electric_field.jl:2-43,
// src/particles/time_marching.jl:66-150).  Built from the same kernels as the full-order path: the only
// new device work is dense -- X = Psi Z for all samples at once (DMMA tile kernel) and three small products.
// ============================================================================================

struct rbm_reduced {
    rbm_pic *pic = nullptr;
    int64_t N = 0;
    int kp = 0, ke = 0;
    int64_t ldN = 0;          // leading dimensions rounded up to even (16-byte aligned columns)
    int kpl = 0, kel = 0;
    DevBuf Psi;               // N x kp (dense copy, ld = ldN)
    DevBuf PPsi;              // ke x kp  = Pi' Psi_x (ld = kel)
    DevBuf Q;                 // kp x ke  = Psi_x' Psi_e (Pi' Psi_e)^-1 (ld = kpl)
    DevBuf z0;                // [2][kp]  Psi'x0, Psi'v0
    DevBuf c_Psi, c_PPsi, c_Q;  // column-pointer arrays
    bool vec_Psi = false, vec_PPsi = false, vec_Q = false;
    int64_t row0 = 0;         // global number of this rank's first row (row-sharded basis, communicator attached)
    // workspaces of rbm_reduced_run, kept between calls (grown on demand): a sweep driven sample by sample, or a host
    // that calls the run once per test parameter, does not pay ~20 cudaMalloc/cudaFree per call
    DevBuf wZx, wZv, wZa, wY, wE, wX, wV, wPar, wPart, wKm, wCoef, wScr, wcZx, wcZv, wcE, wPhi, wWKM, wZxo, wZvo;
};

namespace {

// out[a, j] = A[idx[a] - row0, j] for the rows this rank owns (0 <= idx[a] - row0 < nloc), zero otherwise
__global__ void k_gather_rows(const double *A, int64_t lda, const int64_t *idx, int nrows, int ncols, double *out, int ldo,
                              int64_t row0, int64_t nloc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows * ncols) return;
    const int a = i % nrows, j = i / nrows;
    const int64_t r = idx[a] - row0;
    out[(size_t)j * ldo + a] = (r >= 0 && r < nloc) ? A[(size_t)j * lda + r] : 0.0;
}

// Z[:, s] = z0 for every sample
__global__ void k_fill_cols(double *Z, int ld, const double *z0, int k, int S)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k * S) Z[(size_t)(i / k) * ld + i % k] = z0[i % k];
}

// dst[:, s] += scale[s] * src[:, s]  (product and sum rounded separately, like `z .+= 0.5 .* dt .* zv`)
__global__ void k_axpy_cols(double *dst, const double *src, int ld, const double *scale, int k, int S)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k * S) return;
    const size_t q = (size_t)(i / k) * ld + i % k;
    dst[q] = __dadd_rn(dst[q], __dmul_rn(scale[i / k], src[q]));
}

// The DEIM side of one reduced step is two small dense products per sample around a pointwise field evaluation
// (electric_field.jl:27-31): y = (Pi'Psi) z_x ; e = phi'(y) ; z_a = (Psi'P_e) e, then the kick and the second half drift
// (time_marching.jl:143,146).  As tile GEMMs these were seven launches of a few microseconds of work each (two products, two
// split-K reductions, the gather, two axpys); here they are two kernels.  Block = 32 rows x 8 slices of the contraction,
// a group of RM_SG samples per block: every thread reads its rows' entries once (coalesced over the rows) for all samples of
// the group, the eight partial sums of a row are combined in a fixed order.
constexpr int RM_SG = 8;

// E[a, s] = phi'( sum_j P[a, j] Zx[j, s] )        P: ke x kp (ld ldp), Zx: kp x S (ld ldz), E: ke x S (ld lde)
__global__ void __launch_bounds__(256) k_rm_field(const double *P, int ldp, int ke, int kp, const double *Zx, int ldz, int S,
                                                  const double *coef, Geom g, double *E, int lde)
{
    extern __shared__ double sh[];
    double *z = sh;                                   // [kp][RM_SG]
    double *red = sh + (size_t)kp * RM_SG;            // [8][32][RM_SG]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s0 = blockIdx.y * RM_SG, ns = min(RM_SG, S - s0);
    for (int q = tid; q < kp * RM_SG; q += 256) {
        const int j = q / RM_SG, u = q - j * RM_SG;
        z[q] = u < ns ? Zx[(size_t)(s0 + u) * ldz + j] : 0.0;
    }
    __syncthreads();
    const int a = blockIdx.x * 32 + lane;
    double acc[RM_SG];
#pragma unroll
    for (int u = 0; u < RM_SG; ++u) acc[u] = 0.0;
    if (a < ke) {
        int j = warp;
        for (; j + 24 < kp; j += 32) {   // four independent loads in flight
            const double p0 = P[(size_t)j * ldp + a], p1 = P[(size_t)(j + 8) * ldp + a], p2 = P[(size_t)(j + 16) * ldp + a],
                         p3 = P[(size_t)(j + 24) * ldp + a];
#pragma unroll
            for (int u = 0; u < RM_SG; ++u) {
                acc[u] = fma(p0, z[j * RM_SG + u], acc[u]);
                acc[u] = fma(p1, z[(j + 8) * RM_SG + u], acc[u]);
                acc[u] = fma(p2, z[(j + 16) * RM_SG + u], acc[u]);
                acc[u] = fma(p3, z[(j + 24) * RM_SG + u], acc[u]);
            }
        }
        for (; j < kp; j += 8) {
            const double p0 = P[(size_t)j * ldp + a];
#pragma unroll
            for (int u = 0; u < RM_SG; ++u) acc[u] = fma(p0, z[j * RM_SG + u], acc[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < RM_SG; ++u) red[(warp * 32 + lane) * RM_SG + u] = acc[u];
    __syncthreads();
    {   // thread = (row r, sample u): combine the eight slices in order, evaluate the field at y
        const int r = tid >> 3, u = tid & 7;
        const int ar = blockIdx.x * 32 + r;
        if (ar < ke && u < ns) {
            double y = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) y += red[(w * 32 + r) * RM_SG + u];
            int c;
            double xi;
            locate(y, g, c, xi);
            const double *cf = coef + ((size_t)(s0 + u) * g.nh + c) * 3;
            E[(size_t)(s0 + u) * lde + ar] = fma(fma(__ldg(cf + 2), xi, __ldg(cf + 1)), xi, __ldg(cf));
        }
    }
}

// z_a = Q e ; z_v += dte z_a ; z_x += hdt z_v        Q: kp x ke (ld ldq), E: ke x S (ld lde), Zx, Zv: kp x S (ld ldz)
__global__ void __launch_bounds__(256) k_rm_accel(const double *Q, int ldq, int kp, int ke, const double *E, int lde, int S,
                                                  const double *hdt, const double *dte, double *Zx, double *Zv, int ldz)
{
    extern __shared__ double sh[];
    double *e = sh;                                   // [ke][RM_SG]
    double *red = sh + (size_t)ke * RM_SG;            // [8][32][RM_SG]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s0 = blockIdx.y * RM_SG, ns = min(RM_SG, S - s0);
    for (int q = tid; q < ke * RM_SG; q += 256) {
        const int j = q / RM_SG, u = q - j * RM_SG;
        e[q] = u < ns ? E[(size_t)(s0 + u) * lde + j] : 0.0;
    }
    __syncthreads();
    const int i = blockIdx.x * 32 + lane;
    double acc[RM_SG];
#pragma unroll
    for (int u = 0; u < RM_SG; ++u) acc[u] = 0.0;
    if (i < kp) {
        int j = warp;
        for (; j + 24 < ke; j += 32) {
            const double p0 = Q[(size_t)j * ldq + i], p1 = Q[(size_t)(j + 8) * ldq + i], p2 = Q[(size_t)(j + 16) * ldq + i],
                         p3 = Q[(size_t)(j + 24) * ldq + i];
#pragma unroll
            for (int u = 0; u < RM_SG; ++u) {
                acc[u] = fma(p0, e[j * RM_SG + u], acc[u]);
                acc[u] = fma(p1, e[(j + 8) * RM_SG + u], acc[u]);
                acc[u] = fma(p2, e[(j + 16) * RM_SG + u], acc[u]);
                acc[u] = fma(p3, e[(j + 24) * RM_SG + u], acc[u]);
            }
        }
        for (; j < ke; j += 8) {
            const double p0 = Q[(size_t)j * ldq + i];
#pragma unroll
            for (int u = 0; u < RM_SG; ++u) acc[u] = fma(p0, e[j * RM_SG + u], acc[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < RM_SG; ++u) red[(warp * 32 + lane) * RM_SG + u] = acc[u];
    __syncthreads();
    {
        const int r = tid >> 3, u = tid & 7;
        const int ir = blockIdx.x * 32 + r;
        if (ir < kp && u < ns) {
            double za = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) za += red[(w * 32 + r) * RM_SG + u];
            const size_t q = (size_t)(s0 + u) * ldz + ir;
            // products and sums rounded separately, like `z .+= dt .* za` (time_marching.jl:143,146)
            const double zv = __dadd_rn(Zv[q], __dmul_rn(dte[s0 + u], za));
            Zv[q] = zv;
            Zx[q] = __dadd_rn(Zx[q], __dmul_rn(hdt[s0 + u], zv));
        }
    }
}

// out[k x ns x nparam] slot ts of samples p0.. <- Z[k x S]
__global__ void k_store_z(const double *Z, int ld, int k, int S, double *out, int ns, int ts, int p0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k * S) return;
    const int s = i / k, j = i % k;
    out[(size_t)k * (ts + (size_t)ns * (p0 + s)) + j] = Z[(size_t)s * ld + j];
}

}  // namespace

extern "C" int rbm_reduced_create(rbm_pic *pic, const double *Psi_p, int64_t ldp, int kp, const double *Psi_e,
                                  int64_t lde, int ke, const int64_t *deim_idx, rbm_reduced **out)
{
    if (!pic || !out) return RBM_ERR_INVALID;
    rbm_ctx *ctx = pic->ctx;
    *out = nullptr;
    const int64_t N = pic->np;
    if (!Psi_p || !Psi_e || !deim_idx || kp < 1 || ke < 1 || ldp < N || lde < N)
        return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_reduced_create: bad argument (DimensionMismatch)");
    if (!pic->has_particles) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_reduced_create: call rbm_pic_set_particles first");
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    // With a communicator the bases are row-sharded like the particles (rank r holds rows row0 .. row0 + N of the global
    // basis) and deim_idx holds GLOBAL row numbers, as rbm_deim returns them.
    int64_t row0 = 0, ntotal = N;
    if (ctx->nranks > 1) {
        DevBuf cnt;
        RBM_CUDA(ctx, cnt.alloc(sizeof(double) * ctx->nranks));
        std::vector<double> h(ctx->nranks, 0.0);
        h[ctx->rank] = (double)N;
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
    if (kp > ntotal || ke > ntotal)
        return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_reduced_create: more basis vectors than rows (DimensionMismatch)");
    std::vector<int64_t> hidx(ke);
    RBM_CUDA(ctx, cudaMemcpy(hidx.data(), deim_idx, sizeof(int64_t) * ke, cudaMemcpyDefault));
    for (int a = 0; a < ke; ++a)
        if (hidx[a] < 0 || hidx[a] >= ntotal) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_reduced_create: DEIM row %lld out of range", (long long)hidx[a]);
    rbm_reduced *rm = new (std::nothrow) rbm_reduced();
    if (!rm) return rbm_fail(ctx, RBM_ERR_NOMEM, "out of host memory");
    {   // dynamic shared memory of the fused DEIM-side kernels: (k x RM_SG + 8 x 32 x RM_SG) doubles
        const size_t need = sizeof(double) * ((size_t)std::max(kp, ke) * RM_SG + 8 * 32 * RM_SG);
        if (need > (size_t)ctx->smem_optin)
            return rbm_fail(ctx, RBM_ERR_UNSUPPORTED, "rbm_reduced_create: k_p = %d / k_e = %d too large for the fused reduced step", kp, ke);
        RBM_CUDA(ctx, cudaFuncSetAttribute(k_rm_field, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
        RBM_CUDA(ctx, cudaFuncSetAttribute(k_rm_accel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    }
    rm->pic = pic; rm->N = N; rm->kp = kp; rm->ke = ke; rm->row0 = row0;
    // even leading dimensions: every operand column of the per-step products is 16-byte aligned (vector copies)
    rm->ldN = N + (N & 1); rm->kpl = kp + (kp & 1); rm->kel = ke + (ke & 1);
    const int64_t ldN = rm->ldN;
    const int kpl = rm->kpl, kel = rm->kel;
    auto body = [&]() -> int {
        DevBuf Pe, dIdx, PPe, PPeInv, PtPe, c_Pe, c_x0, c_v0, c_PtPe, c_Inv;
        RBM_CUDA(ctx, rm->Psi.alloc(sizeof(double) * (size_t)ldN * kp));
        RBM_CUDA(ctx, Pe.alloc(sizeof(double) * (size_t)N * ke));
        RBM_CUDA(ctx, cudaMemcpy2DAsync(rm->Psi.p, sizeof(double) * ldN, Psi_p, sizeof(double) * ldp, sizeof(double) * N, kp, cudaMemcpyDefault, ctx->stream));
        RBM_CUDA(ctx, cudaMemcpy2DAsync(Pe.p, sizeof(double) * N, Psi_e, sizeof(double) * lde, sizeof(double) * N, ke, cudaMemcpyDefault, ctx->stream));
        RBM_CUDA(ctx, dIdx.alloc(sizeof(int64_t) * ke));
        RBM_CUDA(ctx, cudaMemcpyAsync(dIdx.p, hidx.data(), sizeof(int64_t) * ke, cudaMemcpyHostToDevice, ctx->stream));
        RBM_CUDA(ctx, rm->PPsi.alloc(sizeof(double) * (size_t)kel * kp));
        RBM_CUDA(ctx, cudaMemsetAsync(rm->PPsi.p, 0, rm->PPsi.bytes, ctx->stream));
        RBM_CUDA(ctx, PPe.alloc(sizeof(double) * (size_t)ke * ke));
        RBM_CUDA(ctx, PPeInv.alloc(sizeof(double) * (size_t)ke * ke));
        RBM_CUDA(ctx, PtPe.alloc(sizeof(double) * (size_t)kp * ke));
        RBM_CUDA(ctx, rm->Q.alloc(sizeof(double) * (size_t)kpl * ke));
        RBM_CUDA(ctx, cudaMemsetAsync(rm->Q.p, 0, rm->Q.bytes, ctx->stream));
        RBM_CUDA(ctx, rm->z0.alloc(sizeof(double) * 2 * (size_t)kp));
        // Pi'Psi_x and Pi'Psi_e: rows of the bases at the DEIM points (electric_field.jl:40-41)
        // (rows owned by another rank contribute zero; the sum over ranks is the full ke x k matrix on every rank)
        k_gather_rows<<<(ke * kp + 255) / 256, 256, 0, ctx->stream>>>(rm->Psi.as<double>(), ldN, dIdx.as<int64_t>(), ke, kp, rm->PPsi.as<double>(), kel, row0, N);
        RBM_LAUNCH_CHECK(ctx);
        k_gather_rows<<<(ke * ke + 255) / 256, 256, 0, ctx->stream>>>(Pe.as<double>(), N, dIdx.as<int64_t>(), ke, ke, PPe.as<double>(), ke, row0, N);
        RBM_LAUNCH_CHECK(ctx);
        RBM_TRY(rbm_allreduce_sum_f64(ctx, rm->PPsi.as<double>(), (size_t)kel * kp));
        RBM_TRY(rbm_allreduce_sum_f64(ctx, PPe.as<double>(), (size_t)ke * ke));
        RBM_TRY(rbm_invert(ctx, PPe.as<double>(), ke, PPeInv.as<double>()));
        bool v1 = false, v2 = false, v3 = false;
        RBM_TRY(rbm_make_cols(ctx, rm->Psi.as<double>(), ldN, kp, rm->c_Psi, &rm->vec_Psi));
        RBM_TRY(rbm_make_cols(ctx, Pe.as<double>(), N, ke, c_Pe, &v1));
        // Psi_x' Psi_e (kp x ke), then times (Pi'Psi_e)^-1  (electric_field.jl:41)
        RBM_TRY(rbm_run_gemm(ctx, true, rm->c_Psi.as<const double *>(), c_Pe.as<const double *>(), kp, ke, N, rm->vec_Psi && v1, PtPe.as<double>(), kp));
        RBM_TRY(rbm_allreduce_sum_f64(ctx, PtPe.as<double>(), (size_t)kp * ke));   // the contraction runs over the global rows
        RBM_TRY(rbm_make_cols(ctx, PtPe.as<double>(), kp, ke, c_PtPe, &v2));
        RBM_TRY(rbm_make_cols(ctx, PPeInv.as<double>(), ke, ke, c_Inv, &v3));
        RBM_TRY(rbm_run_gemm(ctx, false, c_PtPe.as<const double *>(), c_Inv.as<const double *>(), kp, ke, ke, v2 && v3 && (ke % 2 == 0), rm->Q.as<double>(), kpl));
        {   // cuSOLVER's LU reports info = 0 for some exactly singular inputs: check the operator itself
            std::vector<double> hq((size_t)kp * ke);
            RBM_CUDA(ctx, cudaMemcpyAsync(hq.data(), rm->Q.p, sizeof(double) * hq.size(), cudaMemcpyDeviceToHost, ctx->stream));
            RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            for (double q : hq)
                if (!std::isfinite(q))
                    return rbm_fail(ctx, RBM_ERR_NUMERIC, "rbm_reduced_create: singular DEIM interpolation matrix Pi'Psi_e (SingularException)");
        }
        RBM_TRY(rbm_make_cols(ctx, rm->PPsi.as<double>(), kel, kp, rm->c_PPsi, &rm->vec_PPsi));
        RBM_TRY(rbm_make_cols(ctx, rm->Q.as<double>(), kpl, ke, rm->c_Q, &rm->vec_Q));
        // reduced initial condition z = Psi'x0, Psi'v0  (time_marching.jl:122-123)
        RBM_TRY(rbm_make_cols(ctx, pic->x0.as<double>(), N, 1, c_x0, &v1));
        RBM_TRY(rbm_make_cols(ctx, pic->v0.as<double>(), N, 1, c_v0, &v2));
        RBM_TRY(rbm_run_gemm(ctx, true, rm->c_Psi.as<const double *>(), c_x0.as<const double *>(), kp, 1, N, rm->vec_Psi && v1, rm->z0.as<double>(), kp));
        RBM_TRY(rbm_run_gemm(ctx, true, rm->c_Psi.as<const double *>(), c_v0.as<const double *>(), kp, 1, N, rm->vec_Psi && v2, rm->z0.as<double>() + kp, kp));
        RBM_TRY(rbm_allreduce_sum_f64(ctx, rm->z0.as<double>(), 2 * (size_t)kp));
        RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return RBM_OK;
    };
    int rc = body();
    if (rc != RBM_OK) {
        delete rm;
        return rc;
    }
    *out = rm;
    return RBM_OK;
}

extern "C" void rbm_reduced_destroy(rbm_reduced *rm)
{
    if (!rm) return;
    cudaSetDevice(rm->pic->ctx->device);
    cudaStreamSynchronize(rm->pic->ctx->stream);
    delete rm;
}

extern "C" int rbm_reduced_run(rbm_reduced *rm, const double *chi, int nparam, double dt, int nt, int ns, double *Zx,
                               double *Zv, double *X, double *V, double *Phi, double *W, double *K, double *M)
{
    if (!rm) return RBM_ERR_INVALID;
    rbm_pic *pic = rm->pic;
    rbm_ctx *ctx = pic->ctx;
    if (!chi || nparam < 1) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_reduced_run: need nparam >= 1 values of chi");
    if (ns < 2 || nt < ns - 1) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_reduced_run: need ns >= 2 and nt >= ns-1");
    if (!(dt > 0.0)) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_reduced_run: dt must be positive");
    for (int p = 0; p < nparam; ++p)
        if (!(chi[p] > 0.0) || !std::isfinite(chi[p])) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_reduced_run: chi[%d] must be positive", p);
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t N = rm->N;
    const int kp = rm->kp, ke = rm->ke, nh = pic->g.nh;
    const int kpl = rm->kpl, kel = rm->kel;
    const int nsave = nt / (ns - 1);
    const double *w = pic->has_w ? pic->w.as<double>() : nullptr;
    // samples in lock-step: bounded by the reconstructed positions/velocities (2 x N x S doubles)
    size_t free_b = 0, total_b = 0;
    RBM_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
    free_b += rm->wX.bytes + rm->wV.bytes;   // what an earlier run already holds counts as available
    int S = (int)std::min<size_t>((size_t)nparam, std::max<size_t>(1, (size_t)(0.6 * (double)free_b) / (sizeof(double) * 2 * (size_t)N)));
    DevBuf &dZx = rm->wZx, &dZv = rm->wZv, &dZa = rm->wZa, &dY = rm->wY, &dE = rm->wE, &Xbuf = rm->wX, &Vbuf = rm->wV,
           &par = rm->wPar, &part = rm->wPart, &part_km = rm->wKm, &coef = rm->wCoef, &scratch = rm->wScr, &c_Zx = rm->wcZx,
           &c_Zv = rm->wcZv, &c_E = rm->wcE, &dPhi = rm->wPhi, &dWKM = rm->wWKM, &dZxo = rm->wZxo, &dZvo = rm->wZvo;
    Chunks ch = plan_chunks(pic, N, S);
    {
        bool any = false;
        auto want = [&](DevBuf &b, size_t bytes) -> cudaError_t {
            if (b.p && b.bytes >= bytes) return cudaSuccess;
            if (!any) cudaStreamSynchronize(ctx->stream);   // an earlier run may still use the old buffers
            any = true;
            return b.alloc(bytes);
        };
        // a smaller last group is cut into more chunks per sample: size the partial buffers for either plan
        size_t items = (size_t)S * ch.nchunks;
        if (nparam % S) items = std::max(items, (size_t)(nparam % S) * plan_chunks(pic, N, nparam % S).nchunks);
        items = std::max(items, (size_t)S * rbm_recon_deposit_max_chunks(ctx));   // fused reconstruct + deposit: one partial per CTA and sample
        RBM_CUDA(ctx, want(dZx, sizeof(double) * (size_t)kpl * S));
        RBM_CUDA(ctx, want(dZv, sizeof(double) * (size_t)kpl * S));
        RBM_CUDA(ctx, want(dZa, sizeof(double) * (size_t)kpl * S));
        RBM_CUDA(ctx, want(dY, sizeof(double) * (size_t)kel * S));
        RBM_CUDA(ctx, want(dE, sizeof(double) * (size_t)kel * S));
        RBM_CUDA(ctx, want(Xbuf, sizeof(double) * (size_t)N * S));
        RBM_CUDA(ctx, want(Vbuf, sizeof(double) * (size_t)N * S));
        RBM_CUDA(ctx, want(par, sizeof(double) * 4 * (size_t)nparam));
        RBM_CUDA(ctx, want(part, sizeof(double) * (pic->ghist ? (size_t)S : items) * nh));
        RBM_CUDA(ctx, want(part_km, sizeof(double) * items * 2));
        RBM_CUDA(ctx, want(coef, sizeof(double) * (size_t)S * nh * 3));
        RBM_CUDA(ctx, want(scratch, sizeof(double) * (size_t)S * (nh + 2)));
        RBM_CUDA(ctx, want(dPhi, sizeof(double) * (size_t)nparam * ns * nh));
        RBM_CUDA(ctx, want(dWKM, sizeof(double) * 3 * (size_t)nparam * ns));
        if (Zx) RBM_CUDA(ctx, want(dZxo, sizeof(double) * (size_t)kp * ns * nparam));
        if (Zv) RBM_CUDA(ctx, want(dZvo, sizeof(double) * (size_t)kp * ns * nparam));
    }
    RBM_CUDA(ctx, cudaMemsetAsync(dZx.p, 0, sizeof(double) * (size_t)kpl * S, ctx->stream));
    RBM_CUDA(ctx, cudaMemsetAsync(dZv.p, 0, sizeof(double) * (size_t)kpl * S, ctx->stream));
    RBM_CUDA(ctx, cudaMemsetAsync(dZa.p, 0, sizeof(double) * (size_t)kpl * S, ctx->stream));
    RBM_CUDA(ctx, cudaMemsetAsync(dY.p, 0, sizeof(double) * (size_t)kel * S, ctx->stream));
    RBM_CUDA(ctx, cudaMemsetAsync(dE.p, 0, sizeof(double) * (size_t)kel * S, ctx->stream));
    RBM_CUDA(ctx, cudaMemsetAsync(dPhi.p, 0, sizeof(double) * (size_t)nparam * ns * nh, ctx->stream));
    RBM_CUDA(ctx, cudaMemsetAsync(dWKM.p, 0, sizeof(double) * 3 * (size_t)nparam * ns, ctx->stream));
    if (Zx) RBM_CUDA(ctx, cudaMemsetAsync(dZxo.p, 0, sizeof(double) * (size_t)kp * ns * nparam, ctx->stream));
    if (Zv) RBM_CUDA(ctx, cudaMemsetAsync(dZvo.p, 0, sizeof(double) * (size_t)kp * ns * nparam, ctx->stream));
    std::vector<double> hp(4 * (size_t)nparam, 0.0);
    for (int p = 0; p < nparam; ++p) {
        hp[p] = chi[p];
        hp[nparam + p] = 0.5 * (dt * chi[p]);          // time_marching.jl:127,137
        hp[2 * (size_t)nparam + p] = dt * chi[p];
    }
    RBM_CUDA(ctx, cudaMemcpyAsync(par.p, hp.data(), sizeof(double) * hp.size(), cudaMemcpyHostToDevice, ctx->stream));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double *d_chi = par.as<double>(), *d_hdt = d_chi + nparam, *d_dte = d_chi + 2 * (size_t)nparam, *d_zero = d_chi + 3 * (size_t)nparam;
    double *Wd = dWKM.as<double>(), *Kd = Wd + (size_t)nparam * ns, *Md = Kd + (size_t)nparam * ns;
    bool vzx = false, vzv = false, ve = false;
    RBM_TRY(rbm_make_cols(ctx, dZx.as<double>(), kpl, S, c_Zx, &vzx));
    RBM_TRY(rbm_make_cols(ctx, dZv.as<double>(), kpl, S, c_Zv, &vzv));
    RBM_TRY(rbm_make_cols(ctx, dE.as<double>(), kel, S, c_E, &ve));
    const bool vecN = rm->vec_Psi && vzx && vzv;          // all leading dimensions are even: vector copies throughout
    const bool vecY = rm->vec_PPsi && vzx, vecA = rm->vec_Q && ve;

    for (int p0 = 0; p0 < nparam; p0 += S) {
        const int ng = std::min(S, nparam - p0);
        if (ng != S) ch = plan_chunks(pic, N, ng);
        const double *chi_g = d_chi + p0, *hdt_g = d_hdt + p0, *dte_g = d_dte + p0;
        const int nz = kp * ng;
        k_fill_cols<<<(nz + 255) / 256, 256, 0, ctx->stream>>>(dZx.as<double>(), kpl, rm->z0.as<double>(), kp, ng);
        k_fill_cols<<<(nz + 255) / 256, 256, 0, ctx->stream>>>(dZv.as<double>(), kpl, rm->z0.as<double>() + kp, kp, ng);
        RBM_LAUNCH_CHECK(ctx);
        // update!(f, Z, w, t): x = Psi Z for all samples, full deposit, Poisson solve  (electric_field.jl:22-25)
        // Steps that do not save never store X: the deposit is the epilogue of the product (recon.cu)
        const bool fuse_recon = vecN && !w && !pic->ghist && getenv_on("RBM_RM_FUSE_RECON", true);
        auto update_field = [&](const double *vbuf, double *pkm, const SolveOut &o) -> int {
            if (!vbuf && fuse_recon) {
                int nck = 0;
                const int rc = rbm_recon_deposit(ctx, pic->g, rm->Psi.as<double>(), rm->ldN, dZx.as<double>(), kpl, N, ng, kp,
                                                 part.as<double>(), &nck);
                if (rc == RBM_OK)
                    return launch_solve(pic, part.as<double>(), nck, nullptr, 0, chi_g, ng, o, scratch.as<double>(), weights_of(pic));
                if (rc != RBM_ERR_UNSUPPORTED) return rc;
            }
            RBM_TRY(rbm_run_gemm(ctx, false, rm->c_Psi.as<const double *>(), c_Zx.as<const double *>(), N, ng, kp, vecN, Xbuf.as<double>(), N, rm->Psi.as<double>(), rm->ldN));
            return launch_deposit_solve(pic, Xbuf.as<double>(), vbuf, w, vbuf ? d_zero : nullptr, N, N, ng, ch, part.as<double>(),
                                        pkm, pkm, ch.nchunks, chi_g, o, scratch.as<double>(), weights_of(pic));
        };
        int ts = 0;
        auto save_solution = [&]() -> int {   // time_marching.jl:81-105
            RBM_TRY(rbm_run_gemm(ctx, false, rm->c_Psi.as<const double *>(), c_Zv.as<const double *>(), N, ng, kp, vecN, Vbuf.as<double>(), N, rm->Psi.as<double>(), rm->ldN));
            SolveOut o;
            o.phi = dPhi.as<double>() + ((size_t)p0 * ns + ts) * nh; o.phi_stride = (int64_t)ns * nh;
            o.W = Wd + (size_t)p0 * ns + ts; o.K = Kd + (size_t)p0 * ns + ts; o.M = Md + (size_t)p0 * ns + ts; o.w_stride = ns;
            RBM_TRY(update_field(Vbuf.as<double>(), part_km.as<double>(), o));
            const size_t width = sizeof(double) * (size_t)N;
            for (int q = 0; q < ng; ++q) {   // one contiguous copy per sample (pitched copies to pageable memory are slow)
                if (X) RBM_CUDA(ctx, cudaMemcpyAsync(X + ((size_t)(p0 + q) * ns + ts) * N, Xbuf.as<double>() + (size_t)q * N, width, cudaMemcpyDefault, ctx->stream));
                if (V) RBM_CUDA(ctx, cudaMemcpyAsync(V + ((size_t)(p0 + q) * ns + ts) * N, Vbuf.as<double>() + (size_t)q * N, width, cudaMemcpyDefault, ctx->stream));
            }
            if (Zx) k_store_z<<<(nz + 255) / 256, 256, 0, ctx->stream>>>(dZx.as<double>(), kpl, kp, ng, dZxo.as<double>(), ns, ts, p0);
            if (Zv) k_store_z<<<(nz + 255) / 256, 256, 0, ctx->stream>>>(dZv.as<double>(), kpl, kp, ng, dZvo.as<double>(), ns, ts, p0);
            RBM_LAUNCH_CHECK(ctx);
            ++ts;
            return RBM_OK;
        };
        RBM_TRY(save_solution());
        // one time step (time_marching.jl:137-146); identical launches every step, so the second one is captured
        // into a CUDA graph and replayed (the first runs eagerly and sizes the split-K workspace)
        auto step_body = [&]() -> int {
            k_axpy_cols<<<(nz + 255) / 256, 256, 0, ctx->stream>>>(dZx.as<double>(), dZv.as<double>(), kpl, hdt_g, kp, ng);   // :137
            RBM_LAUNCH_CHECK(ctx);
            SolveOut o;
            o.coef = coef.as<double>();
            RBM_TRY(update_field(nullptr, nullptr, o));                                                                // :140 update!
            // efield!: y = (Pi'Psi) z_x ; e = phi'(y) ; z_a = (Psi'P_e) e   (electric_field.jl:27-31), then the kick and the
            // second half drift (:143, :146): two fused kernels
            static const bool fused_small = getenv_on("RBM_RM_FUSED", true);
            if (fused_small) {
                const dim3 g1((unsigned)((ke + 31) / 32), (unsigned)((ng + RM_SG - 1) / RM_SG)), g2((unsigned)((kp + 31) / 32), g1.y);
                {
                    ProfScope ps(ctx, "k_rm_field");
                    k_rm_field<<<g1, 256, sizeof(double) * ((size_t)kp * RM_SG + 8 * 32 * RM_SG), ctx->stream>>>(
                        rm->PPsi.as<double>(), kel, ke, kp, dZx.as<double>(), kpl, ng, coef.as<double>(), pic->g, dE.as<double>(), kel);
                }
                RBM_LAUNCH_CHECK(ctx);
                {
                    ProfScope ps(ctx, "k_rm_accel");
                    k_rm_accel<<<g2, 256, sizeof(double) * ((size_t)ke * RM_SG + 8 * 32 * RM_SG), ctx->stream>>>(
                        rm->Q.as<double>(), kpl, kp, ke, dE.as<double>(), kel, ng, hdt_g, dte_g, dZx.as<double>(), dZv.as<double>(), kpl);
                }
                RBM_LAUNCH_CHECK(ctx);
                return RBM_OK;
            }
            RBM_TRY(rbm_run_gemm(ctx, false, rm->c_PPsi.as<const double *>(), c_Zx.as<const double *>(), ke, ng, kp, vecY, dY.as<double>(), kel, rm->PPsi.as<double>(), kel));
            GatherArgs ga{};
            ga.x = dY.as<double>(); ga.coef = coef.as<double>(); ga.A = dE.as<double>(); ga.snap_stride = kel;
            ga.np = ke; ga.ldx = kel; ga.nsamp = ng; ga.g = pic->g;
            dim3 grid((unsigned)((ke + 255) / 256), ng);
            k_gather<<<grid, 256, 0, ctx->stream>>>(ga);
            RBM_LAUNCH_CHECK(ctx);
            RBM_TRY(rbm_run_gemm(ctx, false, rm->c_Q.as<const double *>(), c_E.as<const double *>(), kp, ng, ke, vecA, dZa.as<double>(), kpl, rm->Q.as<double>(), kpl));
            k_axpy_cols<<<(nz + 255) / 256, 256, 0, ctx->stream>>>(dZv.as<double>(), dZa.as<double>(), kpl, dte_g, kp, ng);   // :143
            k_axpy_cols<<<(nz + 255) / 256, 256, 0, ctx->stream>>>(dZx.as<double>(), dZv.as<double>(), kpl, hdt_g, kp, ng);   // :146
            RBM_LAUNCH_CHECK(ctx);
            return RBM_OK;
        };
        const bool use_graph = !ctx->profiling && getenv_on("RBM_GRAPH", true);
        cudaGraphExec_t step_exec = nullptr;
        int64_t step_launches = 0;   // kernels per step, counted while capturing
        int rc_loop = RBM_OK;
        for (int it = 1; it <= nt && rc_loop == RBM_OK; ++it) {
            if (use_graph && it == 2 && nt > 3) {
                cudaGraph_t graph = nullptr;
                if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                    const int64_t before = ctx->launches;
                    const int rc_cap = step_body();
                    step_launches = ctx->launches - before;
                    ctx->launches = before;              // captured, not executed
                    const cudaError_t e_end = cudaStreamEndCapture(ctx->stream, &graph);
                    if (rc_cap == RBM_OK && e_end == cudaSuccess && graph) {
                        if (cudaGraphInstantiate(&step_exec, graph, 0) != cudaSuccess) step_exec = nullptr;
                    }
                    if (graph) cudaGraphDestroy(graph);
                    (void)cudaGetLastError();
                }
            }
            if (step_exec && it >= 2) {
                if (cudaGraphLaunch(step_exec, ctx->stream) != cudaSuccess)
                    rc_loop = rbm_fail(ctx, RBM_ERR_CUDA, "rbm_reduced_run: graph launch failed");
                else
                    ctx->launches += step_launches;
            } else {
                rc_loop = step_body();
            }
            if (rc_loop == RBM_OK && it % nsave == 0 && ts < ns) rc_loop = save_solution();
        }
        if (step_exec) cudaGraphExecDestroy(step_exec);
        RBM_TRY(rc_loop);
    }
    if (Zx) RBM_CUDA(ctx, cudaMemcpyAsync(Zx, dZxo.p, sizeof(double) * (size_t)kp * ns * nparam, cudaMemcpyDefault, ctx->stream));
    if (Zv) RBM_CUDA(ctx, cudaMemcpyAsync(Zv, dZvo.p, sizeof(double) * (size_t)kp * ns * nparam, cudaMemcpyDefault, ctx->stream));
    if (Phi) RBM_CUDA(ctx, cudaMemcpyAsync(Phi, dPhi.p, sizeof(double) * (size_t)nparam * ns * nh, cudaMemcpyDefault, ctx->stream));
    if (W) RBM_CUDA(ctx, cudaMemcpyAsync(W, Wd, sizeof(double) * (size_t)nparam * ns, cudaMemcpyDefault, ctx->stream));
    if (K) RBM_CUDA(ctx, cudaMemcpyAsync(K, Kd, sizeof(double) * (size_t)nparam * ns, cudaMemcpyDefault, ctx->stream));
    if (M) RBM_CUDA(ctx, cudaMemcpyAsync(M, Md, sizeof(double) * (size_t)nparam * ns, cudaMemcpyDefault, ctx->stream));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RBM_OK;
}

#ifdef RBM_TIMING
// debug build only: per block of the last step-kernel launch: smid, ns at wait release / stream start / stream end / publish
// end (globaltimer), stream cycles
extern "C" int rbm_debug_blocks(unsigned long long *out, int nblocks)
{
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out, rbm::rbm_tblk, sizeof(unsigned long long) * 6 * (size_t)(nblocks < 256 ? nblocks : 256)) ==
                   cudaSuccess ? 0 : -1;
}
// debug build only: out[0] = launches, out[i] = sum of cycles between marks i-1 and i (3 <= i <= 10); clears the counters
extern "C" int rbm_debug_timing(double *out)
{
    unsigned long long acc[16], zero[16] = {0};
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(acc, rbm::rbm_tacc, sizeof(acc));
    for (int i = 0; i < 16; ++i) out[i] = (double)acc[i];
    cudaMemcpyToSymbol(rbm::rbm_tacc, zero, sizeof(acc));
    return 0;
}
#endif
