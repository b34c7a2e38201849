// pic_kernels.cuh -- FP64 sm_100a kernels of the full-order 1D1V Vlasov-Poisson PIC step.
//
// What they replace (reference call sites; the arithmetic itself lives in the
// un-vendored VlasovMethods.jl / PoissonSolvers.jl):
//   locate        cell/offset of a particle in the periodic cubic B-spline grid
//   k_step        x += dt/2 v; a = E(x); v += dt a; x += dt/2 v   (src/particles/time_marching.jl:137-146)
//                 fused with the NEXT step's half drift + charge deposition
//   k_deposit     rhs_j = sum_p w_p B_j(x_p)        (rhs_particles_PBSBasis, time_marching.jl:3)
//   k_solve       mean removal, circulant stiffness solve, phi/chi^2, field energy
//                 (PoissonSolverPBSplines.solve! + ScaledPoissonField, scripts/bump_on_tail_1_training.jl:73,94)
//   k_gather      a_p = phi'(x_p)                   (eval_deriv_PBSBasis, time_marching.jl:3)
//
// Deposition design: FP64 shared-memory atomics are a CAS spin loop on sm_100a
// (ATOMS.CAST.SPIN.64), so every THREAD owns a private histogram hist[bin][tid] in
// shared memory.  With blockDim a multiple of 16 the 64-bit bank of hist[bin][tid] is
// tid%16 for every bin: conflict-free plain LDS/DADD/STS, no atomics, and a fixed
// summation order (bit-reproducible run to run).  The histogram has three alias rows
// (bins nh..nh+2 == bins 0..2) so the four taps of a particle are four constant offsets from
// one base address, with no wrap-around arithmetic.  Blocks publish per-chunk partial sums that
// k_solve adds in a fixed order.
//
// Instruction budget: at the HBM roofline an SM has ~1.3 cycles per particle, i.e. ~85 FP64-pipe
// slots and ~170 issue slots.  The round-1 profile (profiles/) showed the first version at 208
// instructions per particle, issue- and latency-bound; everything below is written to keep the
// per-particle path branch-free and short (floor by magic-number addition, division by Markstein's
// correction of a reciprocal multiply, periodic wrap by one FMA).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace rbm {

struct Geom {
    double L, h, invL, invh;  // invh = RN(1/h), invL = RN(1/L)
    int nh;
    int xhi_max;  // high word of the largest |x| the fast locate accepts (0: always take the exact path)
    int degree;   // spline degree 1, 2 or 3.  The tuned shared-memory kernels are cubic only; degrees 1 and 2 run on the
                  // generic kernels (k_step_global, k_deposit_global, k_gather, k_solve) with zero weights on the unused taps
};

// one thread asks the memory system to pull `bytes` (multiple of 16, 16-byte aligned) into L2
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// 16-byte load of particle state that is read exactly once per step: do not allocate the line in L1
// (the L1 data array is the same SRAM as the shared-memory histograms)
__device__ __forceinline__ double2 ldg_stream(const double *p)
{
#ifdef RBM_EXP_PLAIN_LDG
    return *reinterpret_cast<const double2 *>(p);
#else
    double2 r;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
#endif
}

// fixed-order tree sums of n <= 8 / 16 strided values (missing ones are zero)
__device__ __forceinline__ double tree8(const double *p, const int stride, const int n)
{
    double t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) t[u] = u < n ? p[u * stride] : 0.0;
    return ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
}

__device__ __forceinline__ double tree16(const double *p, const int stride, const int n)
{
    double t[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) t[u] = u < n ? p[u * stride] : 0.0;
#pragma unroll
    for (int w = 1; w < 16; w *= 2)
#pragma unroll
        for (int u = 0; u + w < 16; u += 2 * w) t[u] += t[u + w];
    return t[0];
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Exact restatement of the oracle's locate (Julia mod = fmod + sign fix-up, IEEE division, floor,
// clamp).  Taken only for the rare inputs the fast path cannot certify.  Returns by value so that
// no caller variable has its address taken (keeps the hot path free of local-memory traffic).
struct LocRes {
    double xi;
    int c;
};
static __device__ __noinline__ LocRes locate_exact(double x, double L, double h, int nh)
{
    double r = fmod(x, L);
    if (r < 0.0) r = __dadd_rn(r, L);
    if (r == 0.0) r = 0.0;
    double s = __ddiv_rn(r, h);
    double fl = floor(s);
    int ci = (s == s) ? (int)fl : 0;  // NaN input: keep the index in range
    if (ci > nh - 1) ci = nh - 1;
    if (ci < 0) ci = 0;
    LocRes o;
    o.c = ci;
    o.xi = __dadd_rn(s, -(double)ci);
    return o;
}

// s = mod(x,L)/h ; c = floor(s) ; xi = s - c, bit-identical to locate_exact whenever ok:
//  * q = floor(x/L) by magic-number rounding of fma(x, 1/L, -1/2); with the true q one FMA gives
//    m = x - q L rounded once, which IS Julia's mod(x, L) (fmod is exact; for x < 0 Julia rounds
//    fmod + L once, the same real number).  A q that is off by one makes m < 0 or m >= L;
//  * s = m/h correctly rounded by Markstein's correction of m*RN(1/h) (3 FP64 ops, no MUFU);
//  * floor(s) by magic-number rounding to nearest and a sign correction done on the integer pipe.
// ok is false (-> locate_exact) when m >= L, the cell is out of range, or |x| is too large for
// the magic-number range: a ~1e-15 fraction of random inputs.
__device__ __forceinline__ bool locate_fast(double x, const Geom &g, int &c, double &xi)
{
    const double M = 6755399441055744.0;  // 1.5 * 2^52
    double y = __fma_rn(x, g.invL, -0.5);
    double q = __dadd_rn(__dadd_rn(y, M), -M);
    double m = __fma_rn(-q, g.L, x);
    double q0 = __dmul_rn(m, g.invh);
    double e = __fma_rn(-g.h, q0, m);
    double s = __fma_rn(e, g.invh, q0);
    double t = __dadd_rn(s, M);
    double r = __dadd_rn(t, -M);
    double xi0 = __dadd_rn(s, -r);                 // in [-1/2, 1/2], exact
    int neg = __double2hiint(xi0) >> 31;           // -1 when s is below the nearest integer
    c = __double2loint(t) + neg;
    xi = __dadd_rn(xi0, __hiloint2double(neg & 0x3FF00000, 0));   // + 1.0 when negative (exact)
    bool ok = (m < g.L) & ((unsigned)c < (unsigned)g.nh) & ((__double2hiint(x) & 0x7FFFFFFF) < g.xhi_max);
    return ok;
}

__device__ __forceinline__ void locate(double x, const Geom &g, int &c, double &xi)
{
    if (!locate_fast(x, g, c, xi)) {
        LocRes o = locate_exact(x, g.L, g.h, g.nh);
        c = o.c;
        xi = o.xi;
    }
}

// 6*B_m(xi), m = 0..3 (the 1/6 and a uniform weight are applied once in k_solve);
// B_0(xi) = B_3(1-xi), B_1(xi) = B_2(1-xi)
__device__ __forceinline__ void bspline3x6(double xi, double &b0, double &b1, double &b2, double &b3)
{
    double u = 1.0 - xi;
    double x2 = xi * xi, u2 = u * u;
    b3 = x2 * xi;
    b0 = u2 * u;
    b1 = fma(x2, fma(3.0, xi, -6.0), 4.0);
    b2 = fma(u2, fma(3.0, u, -6.0), 4.0);
}

// 6*B_m(xi) for degree 1, 2 or 3 on the taps m = 0..3 of cells c..c+3 (tap 0 = the particle's own cell; unused taps get
// weight zero).  Degree 2: B = [(1-xi)^2, -2xi^2+2xi+1, xi^2]/2; degree 1: B = [1-xi, xi].  Not used by the tuned kernels.
__device__ __forceinline__ void bsplinex6(double xi, int degree, double &b0, double &b1, double &b2, double &b3)
{
    if (degree == 3) {
        bspline3x6(xi, b0, b1, b2, b3);
    } else if (degree == 2) {
        const double u = 1.0 - xi;
        b0 = 3.0 * (u * u);
        b1 = fma(fma(-6.0, xi, 6.0), xi, 3.0);
        b2 = 3.0 * (xi * xi);
        b3 = 0.0;
    } else {
        b0 = 6.0 * (1.0 - xi);
        b1 = 6.0 * xi;
        b2 = 0.0;
        b3 = 0.0;
    }
}

// per-cell polynomial of phi'(x) = al + xi (be + xi ga) in cell c from the coefficients p0..p3 of taps c..c+3
__device__ __forceinline__ void field_poly(int degree, double ih, double p0, double p1, double p2, double p3, double &al,
                                           double &be, double &ga)
{
    if (degree == 3) {
        al = 0.5 * (p2 - p0) * ih;
        be = ((p0 - 2.0 * p1) + p2) * ih;
        ga = 0.5 * ((p3 - p0) + 3.0 * (p1 - p2)) * ih;
    } else if (degree == 2) {
        al = (p1 - p0) * ih;
        be = ((p0 - 2.0 * p1) + p2) * ih;
        ga = 0.0;
    } else {
        al = (p1 - p0) * ih;
        be = 0.0;
        ga = 0.0;
    }
}

// ---------------------------------------------------------------------------------
// Thread-private histogram in shared memory, hist[bin*NC + column], bins 0..nh+2 (3 alias rows), NC columns.
//  * PAIRED = false: NC = T, every thread owns a column (64-bit bank = tid % 16 for every bin: conflict-free).
//  * PAIRED = true:  NC = T/2, lanes l and l+16 of a warp share column 16*warp + l%16.  The two half-warps update it
//    one after the other (lanes 0-15, __syncwarp, lanes 16-31, __syncwarp): still plain LDS/DADD/STS with a fixed order,
//    the same number of shared-memory wavefronts (a 64-bit access is served per half-warp anyway), half the shared
//    memory -- which is what lets a saved step carry a SECOND histogram at full occupancy.
template <int NC, bool PAIRED> struct SmemHist {
    double *h;  // already offset by this thread's column
    __device__ static __forceinline__ int column(int tid) { return PAIRED ? ((tid >> 5) << 4) + (tid & 15) : tid; }
    __device__ __forceinline__ void rmw(int c, double b0, double b1, double b2, double b3)
    {
        double *p = h + c * NC;
        double h0 = p[0], h1 = p[NC], h2 = p[2 * NC], h3 = p[3 * NC];  // four distinct rows
        p[0] = h0 + b0;
        p[NC] = h1 + b1;
        p[2 * NC] = h2 + b2;
        p[3 * NC] = h3 + b3;
    }
    template <int N> __device__ __forceinline__ void add_n(const int (&c)[N], const double (&b)[N][4])
    {
        if (!PAIRED) {
#pragma unroll
            for (int i = 0; i < N; ++i) rmw(c[i], b[i][0], b[i][1], b[i][2], b[i][3]);
        } else {
            const bool first = (threadIdx.x & 16) == 0;
            if (first) {
#pragma unroll
                for (int i = 0; i < N; ++i) rmw(c[i], b[i][0], b[i][1], b[i][2], b[i][3]);
            }
            __syncwarp();
            if (!first) {
#pragma unroll
                for (int i = 0; i < N; ++i) rmw(c[i], b[i][0], b[i][1], b[i][2], b[i][3]);
            }
            __syncwarp();
        }
    }
};

// Fallback for very large nh (histogram does not fit in shared memory): native FP64
// RED.ADD to a per-sample global vector.
struct GlobalHist {
    double *h;  // [nh]
    __device__ __forceinline__ void add4(int c, int nh, double b0, double b1, double b2, double b3)
    {
        int j1 = c + 1, j2 = c + 2, j3 = c + 3;
        if (j1 >= nh) j1 -= nh;
        if (j2 >= nh) j2 -= nh;
        if (j3 >= nh) j3 -= nh;
        atomicAdd(h + c, b0);
        atomicAdd(h + j1, b1);
        atomicAdd(h + j2, b2);
        atomicAdd(h + j3, b3);
    }
};

// ------------------------------------------------------------------ Poisson solve
// Peer exchange of the per-sample deposit sums (multi-rank, particle-chunk sharding), PUSH based: every rank
// writes its vector into slot [buf][its rank] of EVERY rank's buffer over NVLink and then raises the matching
// flag there, so a reader only polls and reads its own memory.  Buffer of rank r, base[r]:
//   [2][XCH_R][XCH_CAPR] doubles | [2][XCH_R] sequence flags (unsigned long long) | status word | epoch word
// A published vector is identified by its absolute sequence number = epoch + seq: `epoch` lives in the rank's own
// device memory and is advanced by k_xch_advance at the end of every time loop (identically on every rank), `seq`
// counts the exchanges inside the loop.  Kernel arguments therefore do not change from run to run, which is what
// lets a captured CUDA graph of the loop be replayed.  The buffer half in use is the parity of the absolute number.
constexpr int XCH_R = 16;       // max ranks
constexpr int XCH_CAPR = 16384; // doubles per (buffer, source rank): 256 samples x 64 bins (4 MB per context in all)
struct PeerXch {
    double *base[XCH_R];
    int nranks;   // 0: not in use
    int self;     // this rank
    const unsigned long long *epoch;   // this rank's epoch word (device memory)
    unsigned long long seq;            // exchange number inside the current time loop (1, 2, ...)
};
__device__ __forceinline__ double *xch_data(double *base, int buf, int src)
{
    return base + ((size_t)buf * XCH_R + src) * XCH_CAPR;
}
__device__ __forceinline__ unsigned long long *xch_flags(double *base)
{
    return reinterpret_cast<unsigned long long *>(base + (size_t)2 * XCH_R * XCH_CAPR);
}

__device__ __forceinline__ unsigned long long *xch_epoch(double *base) { return xch_flags(base) + 2 * XCH_R + 1; }
__device__ __forceinline__ unsigned long long xch_abs_seq(const PeerXch &p) { return __ldcg(p.epoch) + p.seq; }

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

#ifdef RBM_TIMING
// Debug build only (make EXTRA=-DRBM_TIMING, tools/time_marks.py): cycle stamps of one block's thread 0 along the step kernel.
#ifndef RBM_TBLOCK
#define RBM_TBLOCK 0
#endif
static __device__ unsigned long long rbm_tmark[16];
static __device__ unsigned long long rbm_tacc[16];
static __device__ unsigned long long rbm_tblk[256 * 6];   // per block of the last launch: smid, ns at wait release / stream start / stream end / publish end, stream cycles
__device__ __forceinline__ unsigned long long rbm_gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define RBM_BLK(i) do { if (threadIdx.x == 0 && blockIdx.x < 256) rbm_tblk[blockIdx.x * 6 + (i)] = rbm_gtime(); } while (0)
#define RBM_MARK(i) do { if (blockIdx.x == RBM_TBLOCK && threadIdx.x == 0) rbm_tmark[i] = clock64(); } while (0)
#else
#define RBM_MARK(i) do { } while (0)
#define RBM_BLK(i) do { } while (0)
#endif

struct SolveArgs {
    PeerXch peers;          // nranks > 0: sum the ranks' vectors from peer memory instead of `part`
    const double *part;     // [nsamp][nchunks][nh]
    const double *part_km;  // [nsamp][km_chunks][2] or NULL
    int nchunks, km_chunks;
    double scale;     // rhs = scale * sum   (uniform: w/6, weighted: 1/6)
    double km_scale;  // uniform: w, weighted: 1
    const double *pinv;  // [nh] first row of the circulant pseudo-inverse of K_s
    const double *chi;   // [nsamp]
    double *coef;        // [nsamp][nh][3] or NULL
    double *rhs_out;     // [nsamp][nh] or NULL (scaled deposit before mean removal)
    double *phi_out;     // slot of sample 0 or NULL
    int64_t phi_stride;
    double *W_out, *K_out, *M_out;  // slot of sample 0 or NULL
    int64_t w_stride;
    int nh;
    double h;
    unsigned long long xch_timeout_ns;   // peer wait bound (wall clock)
    int degree;       // 0 is read as 3 (cubic)
    // One rank, uniform weights, nh <= 64: the deposit of the previous step summed over its chunks as 64-bit FIXED-POINT
    // integers (acc[s][bin], value = acc * acc_inv), accumulated by the publishing CTAs with integer atomics -- exact and
    // therefore independent of the order of arrival.  NULL: sum the chunk partials `part`.
    const unsigned long long *acc;
    double acc_inv;
    int no_fast;      // 1: the step kernel solves with the generic solve_sample (A/B switch RBM_FAST_SOLVE=0)
};

// Multi-rank runs: wait until every rank's deposit vector of this exchange has landed in OUR buffer; returns the buffer index.
// Bounded in wall-clock time (%globaltimer, ns; a.xch_timeout_ns, default 10 s): a rank that never publishes must not hang
// the GPU.  On expiry the status word is raised; every later wait sees it and gives up at once, so the rest of the time
// loop drains quickly and rbm_pic_run reports RBM_ERR_COMM (the outputs of such a run are garbage and documented as such
// in the header).  The caller follows with a block barrier.
__device__ __forceinline__ int peer_wait(const SolveArgs &a)
{
    double *mine = a.peers.base[a.peers.self];
    const unsigned long long want = xch_abs_seq(a.peers);
    const int xbuf = (int)(want & 1ull);
    if ((int)threadIdx.x < a.peers.nranks) {
        const unsigned long long *flag = xch_flags(mine) + xbuf * XCH_R + threadIdx.x;
        volatile unsigned long long *status = xch_flags(mine) + 2 * XCH_R;
        unsigned long long t0 = 0;
        unsigned spins = 0;
        while (ld_volatile_u64(flag) < want) {
            if ((++spins & 1023u) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                if (*status != 0ull || now - t0 > a.xch_timeout_ns) {
                    *status = 1ull;
                    break;
                }
            }
        }
    }
    return xbuf;
}

// One block per sample: rho = rhs - mean; phi0 = -pinv(K_s) rho; phi = phi0/chi^2;
// per-cell quadratic of phi'(x); W = chi^2/2 phi'K_s phi.
// The block is organised as G = T/nh groups x nh bins: group g sums every G-th chunk partial
// (independent loads, unrolled) and a quarter of the circulant mat-vec; groups are combined in a
// fixed order, so the result is bit-reproducible.
// tab_out != NULL: also write the gather table, replicated `rep` times across banks, to shared memory
// (tab_out[(c*3+k)*rep + r]); used by k_step, which solves its sample's field itself at item start.
template <int T>
__device__ __forceinline__ void solve_sample(const SolveArgs &a, const int s, double *sm, double *tab_out = nullptr,
                                             int rep = 1)
{
    const int nh = a.nh, tid = threadIdx.x;
    double *rho = sm, *phi = sm + nh, *red = sm + 2 * nh;  // red[64]
    double *pin = red + 64;                                // [nh] pseudo-inverse row
    double *grp = pin + nh;                                // [G][nh] group partials
    const int G = T >= nh ? T / nh : 1;
    const int g = tid / nh, jb = tid - g * nh;             // valid when g < G (T >= nh)
    const double chi = a.chi[s];
    for (int j = tid; j < nh; j += T) pin[j] = a.pinv[j];
    if (a.peers.nranks > 0) {
        double *mine = a.peers.base[a.peers.self];
        const int xbuf = peer_wait(a);
        __syncthreads();
        for (int j = tid; j < nh; j += T) {
            double acc = 0.0;   // fixed rank order: every rank computes bit-identical sums
            for (int r = 0; r < a.peers.nranks; ++r) acc += __ldcv(xch_data(mine, xbuf, r) + (size_t)s * nh + j);
            acc *= a.scale;
            rho[j] = acc;
            if (a.rhs_out) a.rhs_out[(size_t)s * nh + j] = acc;
        }
    } else if (T >= nh) {
        if (g < G) {
            // every CTA of the step kernel does this at item start, on the critical path of the step: all of a thread's
            // chunk partials (<= 32 for 148 chunks and G >= 5) are requested at once -- one L2 round trip, not one per
            // group of eight
            double acc[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] = 0.0;
            const double *p = a.part + (size_t)s * a.nchunks * nh + jb;
            int k = g;
            for (; k < a.nchunks; k += 32 * G) {
                double ld[32];
#pragma unroll
                for (int u = 0; u < 32; ++u) ld[u] = k + u * G < a.nchunks ? __ldcg(p + (size_t)(k + u * G) * nh) : 0.0;
#pragma unroll
                for (int u = 0; u < 32; ++u) acc[u & 7] += ld[u];
            }
            grp[g * nh + jb] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
        }
        __syncthreads();
        for (int j = tid; j < nh; j += T) {
            double acc = 0.0;
            for (int q = 0; q < G; ++q) acc += grp[q * nh + j];
            acc *= a.scale;
            rho[j] = acc;
            if (a.rhs_out) a.rhs_out[(size_t)s * nh + j] = acc;
        }
    } else {
        for (int j = tid; j < nh; j += T) {
            double acc = 0.0;
            const double *p = a.part + (size_t)s * a.nchunks * nh + j;
            for (int k = 0; k < a.nchunks; ++k) acc += __ldcg(p + (size_t)k * nh);
            acc *= a.scale;
            rho[j] = acc;
            if (a.rhs_out) a.rhs_out[(size_t)s * nh + j] = acc;
        }
    }
    __syncthreads();
    // mean (fixed order: strided partials, then serial over the block's warps)
    double part = 0.0;
    for (int j = tid; j < nh; j += T) part += rho[j];
    part = warp_sum(part);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    double mean = 0.0;
    for (int wq = 0; wq < (T >> 5); ++wq) mean += red[wq];
    mean /= (double)nh;
    __syncthreads();
    for (int j = tid; j < nh; j += T) rho[j] -= mean;
    __syncthreads();
    const double ichi2 = 1.0 / (chi * chi);
    if (T >= nh) {
        if (g < G) {  // group g handles i = g, g+G, ...
            double acc = 0.0;
            for (int i = g; i < nh; i += G) {
                int d = jb - i;
                if (d < 0) d += nh;
                acc = fma(pin[d], rho[i], acc);
            }
            grp[g * nh + jb] = acc;
        }
        __syncthreads();
        for (int j = tid; j < nh; j += T) {
            double acc = 0.0;
            for (int q = 0; q < G; ++q) acc += grp[q * nh + j];
            double ph = -acc * ichi2;
            phi[j] = ph;
            if (a.phi_out) a.phi_out[(size_t)s * a.phi_stride + j] = ph;
        }
    } else {
        for (int j = tid; j < nh; j += T) {
            double acc = 0.0;
            for (int i = 0; i < nh; ++i) {
                int d = j - i;
                if (d < 0) d += nh;
                acc = fma(pin[d], rho[i], acc);
            }
            double ph = -acc * ichi2;
            phi[j] = ph;
            if (a.phi_out) a.phi_out[(size_t)s * a.phi_stride + j] = ph;
        }
    }
    __syncthreads();
    if (a.coef) {
        const double ih = 1.0 / a.h;
        for (int c = tid; c < nh; c += T) {
            int c1 = c + 1, c2 = c + 2, c3 = c + 3;
            if (c1 >= nh) c1 -= nh;
            if (c2 >= nh) c2 -= nh;
            if (c3 >= nh) c3 -= nh;
            double p0 = phi[c], p1 = phi[c1], p2 = phi[c2], p3 = phi[c3];
            double *o = a.coef + ((size_t)s * nh + c) * 3;
            field_poly(a.degree ? a.degree : 3, ih, p0, p1, p2, p3, o[0], o[1], o[2]);
        }
    }
    if (tab_out) {
        const double ih = 1.0 / a.h;
        for (int c = tid; c < nh; c += T) {
            int c1 = c + 1, c2 = c + 2, c3 = c + 3;
            if (c1 >= nh) c1 -= nh;
            if (c2 >= nh) c2 -= nh;
            if (c3 >= nh) c3 -= nh;
            double p0 = phi[c], p1 = phi[c1], p2 = phi[c2], p3 = phi[c3];
            double al, be, ga;
            field_poly(a.degree ? a.degree : 3, ih, p0, p1, p2, p3, al, be, ga);
            double *o = tab_out + (size_t)c * 3 * rep;
            for (int r = 0; r < rep; ++r) {
                o[r] = al;
                o[rep + r] = be;
                o[2 * rep + r] = ga;
            }
        }
    }
    if (a.W_out) {
        // stiffness band of the periodic B-splines: (1/h) [-1 2 -1], (1/6h) [-1 -2 6 -2 -1], (1/120h) [-1 -24 -15 80 -15 -24 -1]
        const int deg = a.degree ? a.degree : 3;
        const double band3[7] = {-1.0, -24.0, -15.0, 80.0, -15.0, -24.0, -1.0};
        const double band2[7] = {0.0, -1.0, -2.0, 6.0, -2.0, -1.0, 0.0};
        const double band1[7] = {0.0, 0.0, -1.0, 2.0, -1.0, 0.0, 0.0};
        const double *band = deg == 3 ? band3 : deg == 2 ? band2 : band1;
        const double band_den = deg == 3 ? 120.0 : deg == 2 ? 6.0 : 1.0;
        double e = 0.0;
        for (int i = tid; i < nh; i += T) {
            double r = 0.0;
#pragma unroll
            for (int d = -3; d <= 3; ++d) {
                int j = i + d;
                if (j < 0) j += nh;
                if (j >= nh) j -= nh;
                r = fma(band[d + 3], phi[j], r);
            }
            e = fma(phi[i], r, e);
        }
        e = warp_sum(e);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = e;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int wq = 0; wq < (T >> 5); ++wq) tot += red[wq];
            a.W_out[(size_t)s * a.w_stride] = chi * chi * 0.5 * tot / (band_den * a.h);
            if (a.part_km && a.K_out && a.M_out) {
                double ks = 0.0, ms = 0.0;
                const double *q = a.part_km + (size_t)s * a.km_chunks * 2;
                for (int k = 0; k < a.km_chunks; ++k) {
                    ks += __ldcg(q + 2 * k);
                    ms += __ldcg(q + 2 * k + 1);
                }
                a.K_out[(size_t)s * a.w_stride] = 0.5 * a.km_scale * ks;
                a.M_out[(size_t)s * a.w_stride] = a.km_scale * ms;
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// The field solve of the step kernel (nh <= 64, one rank), organised for LATENCY: it runs in every CTA at the start of every
// step kernel, between the previous grid's last store and this grid's first particle, with 12-16 warps on the SM and
// nothing to hide a dependent chain behind (a dependent FP64 add costs ~40 cycles here, a block barrier ~50).  Measured
// with cycle stamps (tools/time_marks.py): the generic solve_sample above took 9-10 k cycles there, 5 us of a 59 us step.
//   * everything that does not depend on the previous grid is done BEFORE griddepcontrol.wait (solve_fast_prepare):
//     pseudo-inverse row into shared memory, chi, 1/chi^2, 1/h, 1/nh, the thread's indices and kernel parameters;
//   * sums are trees, not chains; loops are unrolled with predicates so that all loads of a phase are in flight at once;
//   * five barriers: chunk partials -> group sums | rho, mean (every warp redundantly: no barrier for the mean) |
//     mat-vec partials | phi + per-cell quadratic | bank-replicated table written by all threads, consecutive addresses
//     (the nh owner threads writing `rep` copies each put all lanes into one bank: cells are 3 rep = 48 doubles apart).
// Same mathematics as solve_sample (rho - mean, circulant pseudo-inverse, phi/chi^2, per-cell quadratic); the summation
// orders are its own (fixed: bit-reproducible from run to run).
// scratch (doubles): grp[16*64] | rc[64] | grp2[8*64] | stg[3*64]
constexpr int FAST_SOLVE_SMEM = 16 * 64 + 64 + 8 * 64 + 3 * 64;

template <int T>
__device__ __forceinline__ bool solve_fast_applies(const SolveArgs &a)
{
    return !a.no_fast && a.nh <= 64 && (a.nh & 1) == 0 && T >= a.nh && !a.coef && !a.W_out && (a.degree == 0 || a.degree == 3) &&
           (a.peers.nranks > 0 || a.acc || (a.nchunks <= 2048 && ((uintptr_t)a.part & 15) == 0));
}

// `pin` = nh + 2 doubles of shared memory outside the scratch (the gather table's first entries: dead until the table is
// written): the pseudo-inverse row, 1/chi^2, 1/h.  A block barrier must follow before solve_fast.
template <int T>
__device__ __forceinline__ void solve_fast_prepare(const SolveArgs &a, const int s, double *pin)
{
    const int nh = a.nh, tid = threadIdx.x;
    for (int j = tid; j < nh; j += T) pin[j] = a.pinv[j];
    if (tid == T - 1) {
        const double chi = a.chi[s];
        pin[nh] = 1.0 / (chi * chi);
        pin[nh + 1] = 1.0 / a.h;
    }
}

// `packed` = the thread's launch-invariant indices, computed once before the wait and kept in ONE register across the particle
// loop: g = tid / nh (7 bits), jb = tid % nh (6 bits), number of chunk-partial PAIRS this thread sums (9 bits), Gc = min(T / nh,
// 8) groups of the mat-vec (4 bits), Gc2 = min(2 T / nh, 16) groups of the partial sums (5 bits).
__device__ __forceinline__ unsigned solve_fast_pack(const int T, const int nh, const int nchunks)
{
    const int tid = threadIdx.x, g = tid / nh, jb = tid - g * nh, G = T / nh, Gc = G < 8 ? G : 8;
    const int G2 = 2 * T / nh, Gc2 = G2 < 16 ? G2 : 16;
    const int total2 = nchunks * (nh / 2), stride2 = Gc2 * (nh / 2);
    int nvalid = tid < stride2 && tid < total2 ? (total2 - tid + stride2 - 1) / stride2 : 0;
    if (nvalid > 511) nvalid = 511;   // (never: solve_fast_applies bounds nchunks)
    return (unsigned)g | ((unsigned)jb << 7) | ((unsigned)nvalid << 13) | ((unsigned)Gc << 22) | ((unsigned)Gc2 << 26);
}

__device__ __forceinline__ double2 ldcg2(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

template <int T>
__device__ __forceinline__ void solve_fast(const SolveArgs &a, const int s, const unsigned packed, double *sm,
                                           const double *pin, double *tab_out, const int rep)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const int qg = (int)(packed & 127u), qjb = (int)((packed >> 7) & 63u), nvalid = (int)((packed >> 13) & 511u);
    const int Gc = (int)((packed >> 22) & 15u), Gc2 = (int)(packed >> 26);
    int nh = a.nh;
    asm volatile("" : "+r"(nh));   // one register, not a constant-bank load at every use
    double *grp = sm, *rc = sm + 1024, *grp2 = sm + 1088, *stg = sm + 1600;
    struct {
        int g, jb;
        double scale;
        double *rhs_out, *phi_out;
    } q{qg, qjb, a.scale, a.rhs_out ? a.rhs_out + (size_t)s * nh : nullptr,
        a.phi_out ? a.phi_out + (size_t)s * a.phi_stride : nullptr};
    // ---- 1: chunk partials -> group sums.  part[chunk][bin] is a flat array of bin PAIRS and (g + u Gc2) nh/2 + j = tid +
    // u Gc2 nh/2: the thread sums every (Gc2 nh/2)-th pair from its own index on -- 16-byte loads, all of them (<= 16 per round)
    // in flight at once, four chains per bin.  Addresses are base + u * stride from registers: with the parameter-space
    // expressions spelled out per load the compiler spent 17 instructions and three constant loads on every one of them,
    // 1600 issue cycles per scheduler.
    // (multi-rank runs: the ranks' chunk-summed vectors are in this rank's exchange buffer once their flags are up)
    const bool summed = a.acc != nullptr || a.peers.nranks > 0;
    int xbuf = 0;
    if (a.peers.nranks > 0) xbuf = peer_wait(a);
    if (!summed && nvalid > 0) {
        double2 acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = make_double2(0.0, 0.0);
        unsigned stride = (unsigned)(Gc2 * (nh >> 1));
        asm volatile("" : "+r"(stride));
        const double2 *pp = reinterpret_cast<const double2 *>(a.part + (size_t)s * a.nchunks * nh) + tid;
        for (int base = 0; base < nvalid; base += 16, pp += 16u * stride) {
            double2 ld[16];
            const int left = nvalid - base;
#pragma unroll
            for (int u = 0; u < 16; ++u) ld[u] = u < left ? ldcg2(pp + (unsigned)u * stride) : make_double2(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                acc[u & 3].x += ld[u].x;
                acc[u & 3].y += ld[u].y;
            }
        }
        reinterpret_cast<double2 *>(grp)[tid] =
            make_double2((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x), (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y));
    }
    if (a.acc == nullptr) __syncthreads();
    RBM_MARK(3);
    // ---- 2: rho and its mean, in every warp (identical bits everywhere); warp 0 stores the centred vector
    {
        const int b0 = lane, b1 = lane + 32;
        double r0 = 0.0, r1 = 0.0;
        if (a.peers.nranks > 0) {   // fixed rank order: every rank computes bit-identical sums
            double *mine = a.peers.base[a.peers.self];
            double t0[XCH_R], t1[XCH_R];
#pragma unroll
            for (int r = 0; r < XCH_R; ++r) {
                const double *pr = xch_data(mine, xbuf, r) + (size_t)s * nh;
                t0[r] = r < a.peers.nranks && b0 < nh ? __ldcv(pr + b0) : 0.0;
                t1[r] = r < a.peers.nranks && b1 < nh ? __ldcv(pr + b1) : 0.0;
            }
#pragma unroll
            for (int r = 0; r < XCH_R; ++r) {
                r0 += t0[r];
                r1 += t1[r];
            }
            r0 *= q.scale;
            r1 *= q.scale;
        } else if (a.acc) {   // the chunk sums are there already: one 8-byte load per bin
            const unsigned long long *pa = a.acc + (size_t)s * nh;
            if (b0 < nh) r0 = (double)(long long)__ldcg(pa + b0) * a.acc_inv * q.scale;
            if (b1 < nh) r1 = (double)(long long)__ldcg(pa + b1) * a.acc_inv * q.scale;
        } else {
            if (b0 < nh) r0 = tree16(grp + b0, nh, Gc2) * q.scale;
            if (b1 < nh) r1 = tree16(grp + b1, nh, Gc2) * q.scale;
        }
        const double tot = warp_sum(r0 + r1);
        const double mean = (nh & (nh - 1)) == 0 ? tot * (1.0 / (double)nh) : tot / (double)nh;   // (exact for powers of two)
        if (tid < 32) {
            if (b0 < nh) rc[b0] = r0 - mean;
            if (b1 < nh) rc[b1] = r1 - mean;
            if (q.rhs_out) {
                if (b0 < nh) q.rhs_out[b0] = r0;
                if (b1 < nh) q.rhs_out[b1] = r1;
            }
        }
    }
    __syncthreads();
    RBM_MARK(4);
    // ---- 3: circulant mat-vec, group g takes i = g, g + Gc, ...; two chains
    if (q.g < Gc) {
        double a0 = 0.0, a1 = 0.0;
        int i = q.g;
#pragma unroll 3
        for (; i + Gc < nh; i += 2 * Gc) {
            int d0 = q.jb - i, d1 = q.jb - i - Gc;
            if (d0 < 0) d0 += nh;
            if (d1 < 0) d1 += nh;
            a0 = fma(pin[d0], rc[i], a0);
            a1 = fma(pin[d1], rc[i + Gc], a1);
        }
        if (i < nh) {
            int d0 = q.jb - i;
            if (d0 < 0) d0 += nh;
            a0 = fma(pin[d0], rc[i], a0);
        }
        grp2[q.g * nh + q.jb] = a0 + a1;
    }
    __syncthreads();
    RBM_MARK(5);
    // ---- 4: phi on the four taps of cell c = tid, per-cell quadratic of phi'
    if (tid < nh) {
        int c1 = tid + 1, c2 = tid + 2, c3 = tid + 3;
        if (c1 >= nh) c1 -= nh;
        if (c2 >= nh) c2 -= nh;
        if (c3 >= nh) c3 -= nh;
        const double ichi2 = pin[nh], ih = pin[nh + 1];
        const double p0 = -tree8(grp2 + tid, nh, Gc) * ichi2, p1 = -tree8(grp2 + c1, nh, Gc) * ichi2;
        const double p2 = -tree8(grp2 + c2, nh, Gc) * ichi2, p3 = -tree8(grp2 + c3, nh, Gc) * ichi2;
        if (q.phi_out) q.phi_out[tid] = p0;
        double al, be, ga;
        field_poly(3, ih, p0, p1, p2, p3, al, be, ga);
        stg[tid] = al;
        stg[64 + tid] = be;
        stg[128 + tid] = ga;
    }
    __syncthreads();
    RBM_MARK(6);
    // ---- 5: the table, replicated across banks, consecutive addresses
    const int per_cell = 3 * rep, total = nh * per_cell;
    if (T % per_cell == 0) {   // entry i + m T belongs to cell c + m T / per_cell, same coefficient: one division per thread
        const int c = tid / per_cell, k = (tid - c * per_cell) / rep;
        const double *src = stg + k * 64 + c;
        double *dst = tab_out + tid;
        const int ncell = nh - c;
#pragma unroll 8
        for (int m = 0; m * (T / per_cell) < ncell; ++m) dst[m * T] = src[m * (T / per_cell)];
    } else {
#pragma unroll 4
        for (int i = tid; i < total; i += T) {
            const int c = i / per_cell, k = (i - c * per_cell) / rep;
            tab_out[i] = stg[k * 64 + c];
        }
    }
    RBM_MARK(7);
}

// shared memory needed by solve_sample<T>: rho[nh] | phi[nh] | red[64] | pin[nh] | grp[T]
__host__ __device__ constexpr int solve_smem_doubles(int nh, int T) { return 3 * nh + 64 + T; }

#ifndef RBM_NO_ENTRY_POINTS   // a second translation unit (recon.cu) includes this header for its device functions only
__global__ void __launch_bounds__(256) k_solve(const __grid_constant__ SolveArgs a)
{
    extern __shared__ double sm[];
    solve_sample<256>(a, blockIdx.x, sm);
}
#endif

// Deferred field solves of the saved slots: block q = slot * ng + sample works on partials q (contiguous: [slot][sample][chunk][nh])
// and writes slot `slot0 + slot` of sample `sample` (Phi: [sample][ns][nh], W/K/M: [sample][ns]).
#ifndef RBM_NO_ENTRY_POINTS   // a second translation unit (recon.cu) includes this header for its device functions only
__global__ void __launch_bounds__(256) k_solve_slots(const __grid_constant__ SolveArgs a, int ng, int slot0)
{
    extern __shared__ double sm[];
    const int q = blockIdx.x, s = q % ng, t = q / ng;
    SolveArgs b = a;
    b.part = a.part + (size_t)q * a.nchunks * a.nh;
    b.part_km = a.part_km ? a.part_km + (size_t)q * a.km_chunks * 2 : nullptr;
    b.chi = a.chi + s;
    b.coef = nullptr;
    b.rhs_out = nullptr;
    b.phi_out = a.phi_out ? a.phi_out + (size_t)s * a.phi_stride + (size_t)(slot0 + t) * a.nh : nullptr;
    b.W_out = a.W_out ? a.W_out + (size_t)s * a.w_stride + slot0 + t : nullptr;
    b.K_out = a.K_out ? a.K_out + (size_t)s * a.w_stride + slot0 + t : nullptr;
    b.M_out = a.M_out ? a.M_out + (size_t)s * a.w_stride + slot0 + t : nullptr;
    b.peers.nranks = 0;
    solve_sample<256>(b, 0, sm);
}
#endif

struct StepArgs {
    double *x, *v;        // [nsamp][ldp] particle state, updated in place
    const double *w;      // [np] weights or NULL (uniform)
    const double *coef;   // [nsamp][nh][3] per-cell quadratic of the field for the kick
    double *part;         // [nsamp][nchunks][nh] deposit partial sums of the NEXT half-drifted x
    double *part_km;      // [nsamp][nchunks][2] sum(w v^2), sum(w v) (SAVE only)
    double *X, *V, *A;    // snapshot slot of sample 0 of this group (any may be NULL)
    int64_t snap_stride;  // elements between samples in X, V, A
    const double *hdt;    // [nsamp] 0.5*(dt*chi)
    const double *dte;    // [nsamp] dt*chi
    int64_t np, ldp, chunk_len;
    int nsamp, nchunks;
    int snap_vec;  // 1: snapshot slots are 16-byte aligned for every sample
    int do_mid;    // 0: skip the deposit (last step)
    int reverse;   // 1: traverse each chunk from its end (alternates per step for L2 reuse)
    // fuse_solve: every block solves the Poisson problem of its sample itself at item start, from the
    // deposit partials the PREVIOUS launch published (sv.part, ping-pong with `part`), straight into its
    // shared-memory gather table: no separate solve kernel, no global round trip, no serial tail.
    int fuse_solve;
    SolveArgs sv;  // part/nchunks/scale/pinv/chi of the field this step kicks with
    // Multi-rank runs: the block that publishes the LAST chunk partial of a sample sums that sample's partials
    // (fixed order) into this rank's exchange buffer; the block that completes the last sample raises the
    // sequence flag.  The peers' next k_step reads the buffer over NVLink (SolveArgs::peers).
    int publish;
    int *tickets;  // [nsamp + 1] zero between launches: per-sample arrivals, then completed samples
    PeerXch pub;   // base[self], cap, buf, seq of the vector being produced by THIS launch
    // DUAL instantiations (saved steps): the update!(efield, x_saved) of save_solution (time_marching.jl:95-99) is
    // deposited by this kernel too, into a second thread-private histogram at the positions it has just computed,
    // so a saved step does not re-read x in a separate deposit pass.
    double *part_save;  // [nsamp][nchunks][nh] deposit partial sums at the SAVED positions
    int save_solve;     // 1: the CTA that publishes a sample's last chunk solves the saved field (sv_save, tickets_save)
    int *tickets_save;  // [nsamp] arrival counters, zero between launches
    SolveArgs sv_save;  // Phi/W/K/M slot of this save
    // Fixed-point deposit sums (see SolveArgs::acc): this launch ADDS its chunk sums, rounded to multiples of 1/acc_scale, to
    // acc_out[s][bin] instead of storing `part`, and clears acc_zero[s][:] (the buffer the launch after next adds to).
    unsigned long long *acc_out, *acc_zero;
    double acc_scale;
    int interleave;   // 1: chunk k = tiles k, k + nchunks, ... (see k_step); 0: contiguous chunks of chunk_len particles
    Geom g;
};

// dst[bin] = sum_t hist[bin][t] (+ the alias row for bins 0..2) in a fixed order; hist is zeroed
// for the next item.  ALIAS = number of alias rows (3 for the padded layout).
template <int T, int NC = T>
__device__ __forceinline__ void block_publish_hist(double *hist, int nh, double *dst, const bool zero = true)
{
    // zero = false: the block's last item of the launch -- nobody reads the histogram again, so it is not cleared
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int nwarps = T / 32;
    for (int bin = warp; bin < nh; bin += nwarps) {
        double s = 0.0;
        double *row = hist + bin * NC;
        if (zero) {
#pragma unroll 4
            for (int t = lane; t < NC; t += 32) {
                s += row[t];
                row[t] = 0.0;
            }
        } else {
#pragma unroll 4
            for (int t = lane; t < NC; t += 32) s += row[t];
        }
        if (bin < 3) {
            double *arow = hist + (nh + bin) * NC;
            if (zero) {
#pragma unroll 4
                for (int t = lane; t < NC; t += 32) {
                    s += arow[t];
                    arow[t] = 0.0;
                }
            } else {
#pragma unroll 4
                for (int t = lane; t < NC; t += 32) s += arow[t];
            }
        }
        s = warp_sum(s);
        if (lane == 0) dst[bin] = s;
    }
    __syncthreads();
}

// The same publish organised for latency (the serial tail of every step kernel): a warp takes six rows at a time, all loads
// in flight, tree sums, the six butterflies interleaved; row sums go through `scratch` (nh + 3 doubles of shared memory
// outside the histogram), then dst[bin] = row[bin] (+ alias row).  Fixed order; 0.6 k cycles where the loop above takes 4-6 k.
template <int T, int NC>
__device__ __forceinline__ void block_publish_hist_fast(double *hist, int nh, double *dst, const bool zero, double *scratch,
                                                        unsigned long long *acc = nullptr, const double acc_scale = 0.0)
{
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int nwarps = T / 32, U = 6, NK = (NC + 31) / 32;
    const int nrows = nh + 3;
    for (int row0 = warp; row0 < nrows; row0 += U * nwarps) {
        double sum[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int row = row0 + u * nwarps;
            const bool live = row < nrows;
            double *p = hist + (live ? row : nrows - 1) * NC + lane;
            double v[NK];
#pragma unroll
            for (int k = 0; k < NK; ++k) v[k] = (NC % 32 == 0 || 32 * k + lane < NC) ? p[32 * k] : 0.0;
            if (zero && live) {
#pragma unroll
                for (int k = 0; k < NK; ++k)
                    if (NC % 32 == 0 || 32 * k + lane < NC) p[32 * k] = 0.0;
            }
#pragma unroll
            for (int w = 1; w < NK; w *= 2)
#pragma unroll
                for (int k = 0; k + w < NK; k += 2 * w) v[k] += v[k + w];
            sum[u] = v[0];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < U; ++u) sum[u] += __shfl_xor_sync(0xffffffffu, sum[u], o);
        if (lane == 0) {
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (row0 + u * nwarps < nrows) scratch[row0 + u * nwarps] = sum[u];
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < nh; j += T) {
        const double r = j < 3 ? scratch[j] + scratch[nh + j] : scratch[j];
        if (acc) atomicAdd(acc + j, (unsigned long long)__double2ll_rn(r * acc_scale));
        else dst[j] = r;
    }
    __syncthreads();
}

// ... and without shuffles: the 32 per-lane column sums of every row go to `scratch` (row stride 33: conflict-free both ways),
// then one thread per bin adds the 32 of its row as a tree (the alias rows are folded in before).  Needs nh * 33 doubles of scratch.
template <int T, int NC>
__device__ __forceinline__ void block_publish_hist_smem(double *hist, int nh, double *dst, const bool zero, double *scratch,
                                                        unsigned long long *acc = nullptr, const double acc_scale = 0.0)
{
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int nwarps = T / 32, U = 6, NK = (NC + 31) / 32;
    for (int row0 = warp; row0 < nh; row0 += U * nwarps) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int row = row0 + u * nwarps;
            if (row < nh) {
                double *p = hist + row * NC + lane;
                double v[NK];
#pragma unroll
                for (int k = 0; k < NK; ++k) v[k] = (NC % 32 == 0 || 32 * k + lane < NC) ? p[32 * k] : 0.0;
                if (zero) {
#pragma unroll
                    for (int k = 0; k < NK; ++k)
                        if (NC % 32 == 0 || 32 * k + lane < NC) p[32 * k] = 0.0;
                }
                if (row < 3) {   // bins 0..2: the alias row nh + bin belongs to the same sum
                    double *pa = hist + (nh + row) * NC + lane;
#pragma unroll
                    for (int k = 0; k < NK; ++k)
                        if (NC % 32 == 0 || 32 * k + lane < NC) {
                            v[k] += pa[32 * k];
                            if (zero) pa[32 * k] = 0.0;
                        }
                }
#pragma unroll
                for (int w = 1; w < NK; w *= 2)
#pragma unroll
                    for (int k = 0; k + w < NK; k += 2 * w) v[k] += v[k + w];
                scratch[row * 33 + lane] = v[0];
            }
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < nh; j += T) {
        const double *p = scratch + j * 33;
        const double r = tree16(p, 1, 16) + tree16(p + 16, 1, 16);
        if (acc) atomicAdd(acc + j, (unsigned long long)__double2ll_rn(r * acc_scale));   // (RED: nothing is returned)
        else dst[j] = r;
    }
    __syncthreads();
}

template <int T> __device__ __forceinline__ void block_publish_km(double ksum, double msum, double *red, double *dst)
{
    ksum = warp_sum(ksum);
    msum = warp_sum(msum);
    if ((threadIdx.x & 31) == 0) {
        red[(threadIdx.x >> 5) * 2] = ksum;
        red[(threadIdx.x >> 5) * 2 + 1] = msum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ks = 0.0, ms = 0.0;
#pragma unroll
        for (int wq = 0; wq < T / 32; ++wq) {
            ks += red[wq * 2];
            ms += red[wq * 2 + 1];
        }
        dst[0] = ks;
        dst[1] = ms;
    }
    __syncthreads();
}

// N particles of one thread, advanced phase by phase so that the N independent dependency chains
// interleave: drift, gather, kick, drift (bit-exact products/sums, never contracted), then the next
// step's half drift + deposit into the thread's histogram.
// live = false (only the tail of a chunk, N = 1): a lane without a particle runs the same instruction sequence on
// zeros and contributes exact zeros, so that every lane of a warp reaches the __syncwarp of the paired histogram.
template <int N, int REP, bool SAVE, bool WEIGHTED, bool DUAL, class Hist>
__device__ __forceinline__ void push_n(double (&x)[N], double (&v)[N], const double (&w)[N], double (&acc)[N],
                                       double hdt, double dte, const Geom &g, const double *tab_lane,
                                       Hist &hist, bool do_mid, double &ksum, double &msum, Hist &hist_save,
                                       const bool live = true)
{
    int c[N];
    double xi[N], xp[N];
    bool ok = true;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        xp[i] = __dadd_rn(x[i], __dmul_rn(hdt, v[i]));
        ok &= locate_fast(xp[i], g, c[i], xi[i]);
    }
    if (!ok) {
#pragma unroll
        for (int i = 0; i < N; ++i) locate(xp[i], g, c[i], xi[i]);
    }
    double xm[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double *tp = tab_lane + c[i] * (3 * REP);
        double al = tp[0], be = tp[REP], ga = tp[2 * REP];
        double a = fma(fma(ga, xi[i], be), xi[i], al);
        double vn = __dadd_rn(v[i], __dmul_rn(dte, a));
        double d = __dmul_rn(hdt, vn);
        double xn = __dadd_rn(xp[i], d);
        xm[i] = __dadd_rn(xn, d);
        x[i] = xn;
        v[i] = vn;
        acc[i] = a;
        if (SAVE) {
            double wv = WEIGHTED ? w[i] * vn : vn;
            if (!live) wv = 0.0;
            ksum = fma(wv, vn, ksum);
            msum += wv;
        }
    }
    if (do_mid) {
        ok = true;
#pragma unroll
        for (int i = 0; i < N; ++i) ok &= locate_fast(xm[i], g, c[i], xi[i]);
        if (!ok) {
#pragma unroll
            for (int i = 0; i < N; ++i) locate(xm[i], g, c[i], xi[i]);
        }
        double b[N][4];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            bspline3x6(xi[i], b[i][0], b[i][1], b[i][2], b[i][3]);
            if (WEIGHTED) {
                b[i][0] *= w[i];
                b[i][1] *= w[i];
                b[i][2] *= w[i];
                b[i][3] *= w[i];
            }
            if (!live) b[i][0] = b[i][1] = b[i][2] = b[i][3] = 0.0;
        }
        hist.template add_n<N>(c, b);
    }
    if (DUAL) {
        // deposit at the saved positions x[i] (the update! of save_solution), second histogram
        ok = true;
#pragma unroll
        for (int i = 0; i < N; ++i) ok &= locate_fast(x[i], g, c[i], xi[i]);
        if (!ok) {
#pragma unroll
            for (int i = 0; i < N; ++i) locate(x[i], g, c[i], xi[i]);
        }
        double b[N][4];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            bspline3x6(xi[i], b[i][0], b[i][1], b[i][2], b[i][3]);
            if (WEIGHTED) {
                b[i][0] *= w[i];
                b[i][1] *= w[i];
                b[i][2] *= w[i];
                b[i][3] *= w[i];
            }
            if (!live) b[i][0] = b[i][1] = b[i][2] = b[i][3] = 0.0;
        }
        hist_save.template add_n<N>(c, b);
    }
}

// shared-memory layout: hist[(nh+3)*T] | (DUAL: hist_save[(nh+3)*T]) | tab[nh*3*REP] | red[64]
// NPT = particles per thread and tile (a tile is NPT*T consecutive particles, NPT/2 double2 per array and thread).
// HM (histogram mode): 0 = one thread-private histogram; 1 = two thread-private histograms (saved steps, small nh: both
// fit); 2 = two histograms with paired columns (saved steps; SmemHist<T/2, true>: the shared memory of one thread-private
// histogram); 3 = one histogram with paired columns (half the shared memory: more threads per SM for a given nh).
template <int T, int NPT, int REP, bool SAVE, bool WEIGHTED, int HM>
__global__ void __launch_bounds__(T, 1) k_step(const __grid_constant__ StepArgs a)
{
    constexpr bool DUAL = HM == 1 || HM == 2;
    static_assert(NPT % 4 == 0 && (!DUAL || SAVE), "tile shape");
    constexpr int NV = NPT / 2;   // double2 vectors per array, thread and tile
    constexpr int NC = HM >= 2 ? T / 2 : T;   // histogram columns
    using Hist = SmemHist<NC, (HM >= 2)>;
    extern __shared__ double smem[];
    const int tid = threadIdx.x, nh = a.g.nh;
    constexpr int NHIST = DUAL ? 2 : 1;
    double *histp = smem;
    double *hist2p = smem + (size_t)(nh + 3) * NC;   // DUAL only
    double *tab = smem + (size_t)NHIST * (nh + 3) * NC;
    double *red = tab + nh * 3 * REP;
    const double *tab_lane = tab + (tid & (REP - 1));
    // Programmatic dependent launch: let the next step's grid start its prologue as SMs free up, and
    // do our own prologue (200 KB of histogram zeroing) before waiting for the previous grid's data.
    RBM_MARK(0);
    asm volatile("griddepcontrol.launch_dependents;");
    for (int i = tid; i < NHIST * (nh + 3) * NC; i += T) histp[i] = 0.0;
    // what the first item's field solve needs and the previous grid does not write: before the wait
    const int items = a.nsamp * a.nchunks;
    const bool fast = a.fuse_solve && solve_fast_applies<T>(a.sv);
    const unsigned packed = fast ? solve_fast_pack(T, nh, a.sv.nchunks) : 0u;
    int s = (int)blockIdx.x / a.nchunks, k = (int)blockIdx.x - s * a.nchunks;   // (sample, chunk) of the first item
    if (fast && (int)blockIdx.x < items) solve_fast_prepare<T>(a.sv, s, tab);
    RBM_MARK(1);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    RBM_MARK(2);
    RBM_BLK(1);
#ifdef RBM_TIMING
    if (tid == 0 && blockIdx.x < 256) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        rbm_tblk[blockIdx.x * 6] = smid;
    }
    long long blk_c0 = 0;
#endif
    Hist hist{histp + Hist::column(tid)};
    Hist hist_save{hist2p + Hist::column(tid)};
    const Geom g = a.g;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        if (item != (int)blockIdx.x) {
            s = item / a.nchunks;
            k = item - s * a.nchunks;
            if (fast) solve_fast_prepare<T>(a.sv, s, tab);   // (the table is dead: barrier at the previous item's end)
        }
        const double hdt = a.hdt[s], dte = a.dte[s];   // (requested before the solve: not waited for after it)
        if (a.acc_zero && k == 0) {   // nobody reads or adds to this buffer during this launch
            for (int j = tid; j < nh; j += T) a.acc_zero[(size_t)s * nh + j] = 0ull;
        }
        __syncthreads();
        if (fast) {
            // the (all-zero) histogram doubles as scratch; the region is zeroed again afterwards
            solve_fast<T>(a.sv, s, packed, histp, tab, tab, REP);
            __syncthreads();
            for (int i = tid; i < FAST_SOLVE_SMEM; i += T) histp[i] = 0.0;
        } else if (a.fuse_solve) {
            solve_sample<T>(a.sv, s, histp, tab, REP);
            __syncthreads();
            for (int i = tid; i < solve_smem_doubles(nh, T); i += T) histp[i] = 0.0;
        } else {
            for (int i = tid; i < nh * 3 * REP; i += T) tab[i] = a.coef[(size_t)s * nh * 3 + i / REP];
        }
        __syncthreads();
        RBM_MARK(8);
        RBM_BLK(2);
#ifdef RBM_TIMING
        blk_c0 = clock64();
#endif
        constexpr int64_t tile = NPT * (int64_t)T;
        // The particles of chunk k: a contiguous range (a.interleave == 0), or every nchunks-th tile (interleaved: tile k,
        // k + nchunks, ...; the sample's last partial tile goes to the chunk that has one tile less).  Interleaved, all CTAs
        // of a launch work in the same window of the arrays at any time, so they see the same mix of L2 hits (the part
        // of the state the previous, opposite sweep wrote last) and HBM reads: the CTAs' finishing times, whose
        // maximum every step of the time loop waits for, spread by +-1 % instead of +-5 %.
        int64_t lo, hi, nfull, tstep;   // first particle of the first full tile, [rem_lo = lo + nfull * tstep .. hi) remainder
        if (a.interleave) {
            const int64_t ntile = a.np / tile;
            nfull = k < ntile ? (ntile - k + a.nchunks - 1) / a.nchunks : 0;
            tstep = tile * a.nchunks;
            lo = (int64_t)k * tile;
            hi = (int)(ntile % a.nchunks) == k ? a.np : -1;   // (this chunk's remainder range is [ntile * tile, np))
        } else {
            lo = (int64_t)k * a.chunk_len;
            hi = lo + a.chunk_len;
            if (hi > a.np) hi = a.np;
            nfull = hi > lo ? (hi - lo) / tile : 0;
            tstep = tile;
        }
        const int64_t rem_lo = a.interleave ? (a.np / tile) * tile : lo + nfull * tile;
        double *xs = a.x + (size_t)s * a.ldp, *vs = a.v + (size_t)s * a.ldp;
        double *Xs = a.X ? a.X + (size_t)s * a.snap_stride : nullptr;
        double *Vs = a.V ? a.V + (size_t)s * a.snap_stride : nullptr;
        double *As = a.A ? a.A + (size_t)s * a.snap_stride : nullptr;
        double ksum = 0.0, msum = 0.0;
        const bool do_mid = a.do_mid != 0;
        // Traversal direction alternates from step to step (a.reverse): the particles written last in
        // the previous step are still in the 126 MB L2 and are read first in this one, so a large part
        // of the state never makes the round trip to HBM.  `stride` is +tstep or -tstep.
        const bool rev = a.reverse != 0;
        const int64_t stride = rev ? -tstep : tstep;
        auto remainder = [&]() {   // (< one tile at the end of the chunk), one particle per thread per pass
            // block-uniform trip count: lanes past the end run on zeros (see push_n) and store nothing
            for (int64_t base = rem_lo; base < hi; base += T) {
                const int64_t i = base + tid;
                const bool live = i < hi;
                double px[1] = {live ? xs[i] : 0.0}, pv[1] = {live ? vs[i] : 0.0}, pa[1];
                const double pw[1] = {WEIGHTED && live ? a.w[i] : 0.0};
                push_n<1, REP, SAVE, WEIGHTED, DUAL>(px, pv, pw, pa, hdt, dte, g, tab_lane, hist, do_mid, ksum, msum, hist_save,
                                                          live);
                if (live) {
                    xs[i] = px[0];
                    vs[i] = pv[0];
                    if (SAVE) {
                        if (Xs) Xs[i] = px[0];
                        if (Vs) Vs[i] = pv[0];
                        if (As) As[i] = pa[0];
                    }
                }
            }
        };
        if (rev) remainder();
        // ---- full tiles: NV double2 per array per thread, next tile prefetched into registers,
        //      the tile after that prefetched into L2 ----
        double2 cx[NV], cv[NV], cw[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) cw[j] = make_double2(0.0, 0.0);
        int64_t i0 = lo + (rev ? (nfull - 1) * tstep : 0) + 2 * tid;
        if (nfull > 0) {
#pragma unroll
            for (int j = 0; j < NV; ++j) cx[j] = ldg_stream(xs + i0 + 2 * T * j);
#pragma unroll
            for (int j = 0; j < NV; ++j) cv[j] = ldg_stream(vs + i0 + 2 * T * j);
            if (WEIGHTED) {
#pragma unroll
                for (int j = 0; j < NV; ++j) cw[j] = ldg_stream(a.w + i0 + 2 * T * j);
            }
        }
        for (int64_t t = 0; t < nfull; ++t, i0 += stride) {
            double2 nx[NV], nv[NV], nw[NV];
#pragma unroll
            for (int j = 0; j < NV; ++j) nw[j] = make_double2(0.0, 0.0);
            if (t + 1 < nfull) {
#pragma unroll
                for (int j = 0; j < NV; ++j) nx[j] = ldg_stream(xs + i0 + stride + 2 * T * j);
#pragma unroll
                for (int j = 0; j < NV; ++j) nv[j] = ldg_stream(vs + i0 + stride + 2 * T * j);
                if (WEIGHTED) {
#pragma unroll
                    for (int j = 0; j < NV; ++j) nw[j] = ldg_stream(a.w + i0 + stride + 2 * T * j);
                }
            }
            if (tid == 0 && t + 3 < nfull) {   // one bulk L2 prefetch per tile and array (12 KB each)
                prefetch_l2_bulk(xs + i0 + 3 * stride, (unsigned)(tile * sizeof(double)));
                prefetch_l2_bulk(vs + i0 + 3 * stride, (unsigned)(tile * sizeof(double)));
            }
#pragma unroll
            for (int q = 0; q < NPT / 4; ++q) {   // four particles (two vectors per array) at a time
                double px[4] = {cx[2 * q].x, cx[2 * q].y, cx[2 * q + 1].x, cx[2 * q + 1].y};
                double pv[4] = {cv[2 * q].x, cv[2 * q].y, cv[2 * q + 1].x, cv[2 * q + 1].y};
                const double pw[4] = {cw[2 * q].x, cw[2 * q].y, cw[2 * q + 1].x, cw[2 * q + 1].y};
                double pa[4];
                push_n<4, REP, SAVE, WEIGHTED, DUAL>(px, pv, pw, pa, hdt, dte, g, tab_lane, hist, do_mid, ksum, msum, hist_save);
                const int64_t j0 = i0 + 2 * T * (2 * q), j1 = i0 + 2 * T * (2 * q + 1);
                const double2 ox0 = make_double2(px[0], px[1]), ox1 = make_double2(px[2], px[3]);
                const double2 ov0 = make_double2(pv[0], pv[1]), ov1 = make_double2(pv[2], pv[3]);
                *reinterpret_cast<double2 *>(xs + j0) = ox0;
                *reinterpret_cast<double2 *>(xs + j1) = ox1;
                *reinterpret_cast<double2 *>(vs + j0) = ov0;
                *reinterpret_cast<double2 *>(vs + j1) = ov1;
                if (SAVE) {
                    const double2 oa0 = make_double2(pa[0], pa[1]), oa1 = make_double2(pa[2], pa[3]);
                    if (a.snap_vec) {
                        if (Xs) { __stcs(reinterpret_cast<double2 *>(Xs + j0), ox0); __stcs(reinterpret_cast<double2 *>(Xs + j1), ox1); }
                        if (Vs) { __stcs(reinterpret_cast<double2 *>(Vs + j0), ov0); __stcs(reinterpret_cast<double2 *>(Vs + j1), ov1); }
                        if (As) { __stcs(reinterpret_cast<double2 *>(As + j0), oa0); __stcs(reinterpret_cast<double2 *>(As + j1), oa1); }
                    } else {
                        if (Xs) { Xs[j0] = ox0.x; Xs[j0 + 1] = ox0.y; Xs[j1] = ox1.x; Xs[j1 + 1] = ox1.y; }
                        if (Vs) { Vs[j0] = ov0.x; Vs[j0 + 1] = ov0.y; Vs[j1] = ov1.x; Vs[j1 + 1] = ov1.y; }
                        if (As) { As[j0] = oa0.x; As[j0 + 1] = oa0.y; As[j1] = oa1.x; As[j1 + 1] = oa1.y; }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                cx[j] = nx[j];
                cv[j] = nv[j];
                if (WEIGHTED) cw[j] = nw[j];
            }
        }
        if (!rev) remainder();
        const bool more = item + (int)gridDim.x < items;   // another item follows in this launch: leave the histograms zeroed
        RBM_MARK(9);
        RBM_BLK(3);
#ifdef RBM_TIMING
        if (tid == 0 && blockIdx.x < 256) rbm_tblk[blockIdx.x * 6 + 5] = (unsigned long long)(clock64() - blk_c0);
#endif
        // (row sums go through the gather table's memory: the table is dead once the chunk's particles are pushed)
        unsigned long long *const acc_s = a.acc_out ? a.acc_out + (size_t)s * nh : nullptr;
        if (REP >= 16) {   // (nh * 33 doubles of scratch fit into the table's nh * 3 * REP)
            if (do_mid) block_publish_hist_smem<T, NC>(histp, nh, a.part + ((size_t)s * a.nchunks + k) * nh, more, tab, acc_s, a.acc_scale);
            if (DUAL) block_publish_hist_smem<T, NC>(hist2p, nh, a.part_save + ((size_t)s * a.nchunks + k) * nh, more, tab);
        } else {
            if (do_mid) block_publish_hist_fast<T, NC>(histp, nh, a.part + ((size_t)s * a.nchunks + k) * nh, more, tab, acc_s, a.acc_scale);
            if (DUAL) block_publish_hist_fast<T, NC>(hist2p, nh, a.part_save + ((size_t)s * a.nchunks + k) * nh, more, tab);
        }
        RBM_MARK(10);
        RBM_BLK(4);
#ifdef RBM_TIMING
        if (blockIdx.x == RBM_TBLOCK && tid == 0 && a.fuse_solve && !SAVE) {
            for (int i = 3; i < 11; ++i) rbm_tacc[i] += rbm_tmark[i] - rbm_tmark[i - 1];
            rbm_tacc[0] += 1;
        }
#endif
        if (SAVE) block_publish_km<T>(ksum, msum, red, a.part_km + ((size_t)s * a.nchunks + k) * 2);
        if (DUAL && a.save_solve) {
            // update!(efield, x_saved) of this save: the last chunk of the sample to arrive sums the partials of the
            // second histogram and solves for Phi, W (and K, M from the chunk partials) -- no deposit pass, no solve kernel
            __shared__ int s_last;
            if (tid == 0) {
                __threadfence();
                s_last = atomicAdd(a.tickets_save + s, 1) == a.nchunks - 1;
                if (s_last) a.tickets_save[s] = 0;
            }
            __syncthreads();
            if (s_last) {
                __threadfence();
                solve_sample<T>(a.sv_save, s, histp);   // the all-zero histogram doubles as scratch
                __syncthreads();
                for (int i = tid; i < solve_smem_doubles(nh, T); i += T) histp[i] = 0.0;
            }
            __syncthreads();
        }
        if (do_mid && a.publish) {
            __shared__ int s_role;   // 0: nothing, 1: last chunk of the sample, 2: ... and last sample
            if (tid == 0) {
                __threadfence();
                s_role = (atomicAdd(a.tickets + s, 1) == a.nchunks - 1) ? 1 : 0;
                if (s_role) a.tickets[s] = 0;
            }
            __syncthreads();
            if (s_role) {
                __threadfence();
                const unsigned long long pseq = xch_abs_seq(a.pub);
                const int pbuf = (int)(pseq & 1ull);
                // grp scratch = the first T doubles of the (all-zero) histogram, zeroed again below
                const int G = T >= nh ? T / nh : 1;
                const int gq = tid / nh, jb = tid - gq * nh;
                if (T >= nh) {
                    if (gq < G) {
                        // serial tail of the launch (every other CTA has finished): all of a thread's chunk partials are
                        // requested at once -- one L2 round trip instead of one per group of four
                        const double *p = a.part + (size_t)s * a.nchunks * nh + jb;
                        double ac[4] = {0.0, 0.0, 0.0, 0.0};
                        for (int kk = gq; kk < a.nchunks; kk += 32 * G) {
                            double ld[32];
#pragma unroll
                            for (int u = 0; u < 32; ++u) ld[u] = kk + u * G < a.nchunks ? __ldcg(p + (size_t)(kk + u * G) * nh) : 0.0;
#pragma unroll
                            for (int u = 0; u < 32; ++u) ac[u & 3] += ld[u];
                        }
                        histp[gq * nh + jb] = (ac[0] + ac[1]) + (ac[2] + ac[3]);
                    }
                    __syncthreads();
                    if (tid < nh) {
                        double acc = 0.0;
                        for (int q = 0; q < G; ++q) acc += histp[q * nh + tid];
                        for (int r = 0; r < a.pub.nranks; ++r)   // push to every rank (NVLink stores)
                            xch_data(a.pub.base[r], pbuf, a.pub.self)[(size_t)s * nh + tid] = acc;
                    }
                    __syncthreads();
                    if (gq < G) histp[gq * nh + jb] = 0.0;
                } else {
                    for (int j = tid; j < nh; j += T) {
                        double acc = 0.0;
                        const double *p = a.part + (size_t)s * a.nchunks * nh + j;
                        for (int kk = 0; kk < a.nchunks; ++kk) acc += __ldcg(p + (size_t)kk * nh);
                        for (int r = 0; r < a.pub.nranks; ++r)
                            xch_data(a.pub.base[r], pbuf, a.pub.self)[(size_t)s * nh + j] = acc;
                    }
                }
                __threadfence_system();
                __syncthreads();
                if (tid == 0) {
                    if (atomicAdd(a.tickets + a.nsamp, 1) == a.nsamp - 1) {   // every sample of this launch is in the buffer
                        a.tickets[a.nsamp] = 0;
                        __threadfence_system();
                        for (int r = 0; r < a.pub.nranks; ++r) {
                            volatile unsigned long long *flag = xch_flags(a.pub.base[r]) + pbuf * XCH_R + a.pub.self;
                            *flag = pseq;
                        }
                        __threadfence_system();
                    }
                }
            }
            __syncthreads();
        }
    }
}

// Large-nh fallback (histogram does not fit in shared memory): same step, deposit by native
// RED.ADD.F64 into part[s][:] (zeroed by the host), gather table read through L1.
template <bool SAVE, bool WEIGHTED>
__global__ void __launch_bounds__(256) k_step_global(const __grid_constant__ StepArgs a)
{
    __shared__ double red[32];
    const Geom g = a.g;
    const int tid = threadIdx.x;
    const int items = a.nsamp * a.nchunks;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int s = item / a.nchunks, k = item - s * a.nchunks;
        const double hdt = a.hdt[s], dte = a.dte[s];
        const int64_t lo = (int64_t)k * a.chunk_len;
        int64_t hi = lo + a.chunk_len;
        if (hi > a.np) hi = a.np;
        double *xs = a.x + (size_t)s * a.ldp, *vs = a.v + (size_t)s * a.ldp;
        double *Xs = a.X ? a.X + (size_t)s * a.snap_stride : nullptr;
        double *Vs = a.V ? a.V + (size_t)s * a.snap_stride : nullptr;
        double *As = a.A ? a.A + (size_t)s * a.snap_stride : nullptr;
        const double *cf = a.coef + (size_t)s * g.nh * 3;
        GlobalHist hist{a.part + (size_t)s * g.nh};
        double ksum = 0.0, msum = 0.0;
        for (int64_t i = lo + tid; i < hi; i += blockDim.x) {
            double xv = xs[i], vv = vs[i];
            const double wv = WEIGHTED ? a.w[i] : 0.0;
            double xp = __dadd_rn(xv, __dmul_rn(hdt, vv));
            int c;
            double xi;
            locate(xp, g, c, xi);
            double al = __ldg(cf + c * 3), be = __ldg(cf + c * 3 + 1), ga = __ldg(cf + c * 3 + 2);
            double acc = fma(fma(ga, xi, be), xi, al);
            double vn = __dadd_rn(vv, __dmul_rn(dte, acc));
            double d = __dmul_rn(hdt, vn);
            double xn = __dadd_rn(xp, d);
            if (a.do_mid) {
                locate(__dadd_rn(xn, d), g, c, xi);
                double b0, b1, b2, b3;
                bsplinex6(xi, g.degree, b0, b1, b2, b3);
                if (WEIGHTED) { b0 *= wv; b1 *= wv; b2 *= wv; b3 *= wv; }
                hist.add4(c, g.nh, b0, b1, b2, b3);
            }
            xs[i] = xn;
            vs[i] = vn;
            if (SAVE) {
                double t = WEIGHTED ? wv * vn : vn;
                ksum = fma(t, vn, ksum);
                msum += t;
                if (Xs) Xs[i] = xn;
                if (Vs) Vs[i] = vn;
                if (As) As[i] = acc;
            }
        }
        if (SAVE) block_publish_km<256>(ksum, msum, red, a.part_km + ((size_t)s * a.nchunks + k) * 2);
    }
}

// ------------------------------------------------------------------ deposit only
struct DepositArgs {
    const double *x;      // [nsamp][ldx] positions (ldx = 0: the same array for every sample)
    const double *v;      // optional: deposit x + hdt[s]*v (first half drift), same stride
    const double *w;      // [np] or NULL
    const double *hdt;    // [nsamp] (only with v)
    double *part;         // [nsamp][nchunks][nh]
    double *part_km;      // optional [nsamp][nchunks][2]: sum(w v^2), sum(w v) of v
    int64_t np, ldx, chunk_len;
    int nsamp, nchunks;
    int vec;              // x, v, w and the sample stride are 16-byte aligned: streaming double2 path
    int fuse_solve;       // the CTA that publishes a sample's last chunk also solves its field (sv, tickets)
    int *tickets;         // [nsamp] arrival counters, zero between launches
    SolveArgs sv;
    Geom g;
};

template <int T, bool WEIGHTED>
__global__ void __launch_bounds__(T, 1) k_deposit(const __grid_constant__ DepositArgs a)
{
    extern __shared__ double smem[];
    const int tid = threadIdx.x, nh = a.g.nh;
    double *histp = smem;
    double *red = smem + (size_t)(nh + 3) * T;
    for (int i = tid; i < (nh + 3) * T; i += T) histp[i] = 0.0;
    SmemHist<T, false> hist{histp + tid};
    const Geom g = a.g;
    const int items = a.nsamp * a.nchunks;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int s = item / a.nchunks, k = item - s * a.nchunks;
        const int64_t lo = (int64_t)k * a.chunk_len;
        int64_t hi = lo + a.chunk_len;
        if (hi > a.np) hi = a.np;
        const double *xs = a.x + (size_t)s * a.ldx;
        const double *vs = a.v ? a.v + (size_t)s * a.ldx : nullptr;
        const double hdt = a.v ? a.hdt[s] : 0.0;
        double ksum = 0.0, msum = 0.0;
        // four independent particles: half drift (optional), K/M sums, locate, weights, histogram
        auto deposit4 = [&](double (&xv)[4], const double (&vv)[4], const double (&wv)[4]) {
            double b[4][4], xi[4];
            int c[4];
            bool ok = true;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (vs) {
                    xv[u] = __dadd_rn(xv[u], __dmul_rn(hdt, vv[u]));
                    double t = WEIGHTED ? wv[u] * vv[u] : vv[u];
                    ksum = fma(t, vv[u], ksum);
                    msum += t;
                }
                ok &= locate_fast(xv[u], g, c[u], xi[u]);
            }
            if (!ok) {
#pragma unroll
                for (int u = 0; u < 4; ++u) locate(xv[u], g, c[u], xi[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                bspline3x6(xi[u], b[u][0], b[u][1], b[u][2], b[u][3]);
                if (WEIGHTED) { b[u][0] *= wv[u]; b[u][1] *= wv[u]; b[u][2] *= wv[u]; b[u][3] *= wv[u]; }
            }
            hist.template add_n<4>(c, b);
        };
        int64_t i = lo + tid;
        if (a.vec) {
            // streaming layout of k_step: two double2 per array and thread, the next tile prefetched into registers,
            // the tile after the next two into L2 by one bulk prefetch per array
            constexpr int64_t tile = 4 * (int64_t)T;
            const int64_t nfull = (hi - lo) / tile;
            int64_t i0 = lo + 2 * tid;
            double2 cx0, cx1, cv0, cv1, cw0, cw1;
            cv0 = cv1 = cw0 = cw1 = make_double2(0.0, 0.0);
            if (nfull > 0) {
                cx0 = ldg_stream(xs + i0); cx1 = ldg_stream(xs + i0 + 2 * T);
                if (vs) { cv0 = ldg_stream(vs + i0); cv1 = ldg_stream(vs + i0 + 2 * T); }
                if (WEIGHTED) { cw0 = ldg_stream(a.w + i0); cw1 = ldg_stream(a.w + i0 + 2 * T); }
            }
            for (int64_t t = 0; t < nfull; ++t, i0 += tile) {
                double2 nx0, nx1, nv0, nv1, nw0, nw1;
                nv0 = nv1 = nw0 = nw1 = make_double2(0.0, 0.0);
                if (t + 1 < nfull) {
                    nx0 = ldg_stream(xs + i0 + tile); nx1 = ldg_stream(xs + i0 + tile + 2 * T);
                    if (vs) { nv0 = ldg_stream(vs + i0 + tile); nv1 = ldg_stream(vs + i0 + tile + 2 * T); }
                    if (WEIGHTED) { nw0 = ldg_stream(a.w + i0 + tile); nw1 = ldg_stream(a.w + i0 + tile + 2 * T); }
                }
                if (tid == 0 && t + 3 < nfull) {
                    prefetch_l2_bulk(xs + i0 - 2 * tid + 3 * tile, (unsigned)(tile * sizeof(double)));
                    if (vs) prefetch_l2_bulk(vs + i0 - 2 * tid + 3 * tile, (unsigned)(tile * sizeof(double)));
                }
                double xv[4] = {cx0.x, cx0.y, cx1.x, cx1.y};
                const double vv[4] = {cv0.x, cv0.y, cv1.x, cv1.y}, wv[4] = {cw0.x, cw0.y, cw1.x, cw1.y};
                deposit4(xv, vv, wv);
                cx0 = nx0; cx1 = nx1; cv0 = nv0; cv1 = nv1; cw0 = nw0; cw1 = nw1;
            }
            i = lo + nfull * tile + tid;
        } else {
            // unaligned arrays: the same particle-to-thread assignment with scalar loads, so that the result does
            // not depend on the alignment of the caller's buffers (bit-identical histograms)
            constexpr int64_t tile = 4 * (int64_t)T;
            const int64_t nfull = (hi - lo) / tile;
            for (int64_t t = 0; t < nfull; ++t) {
                const int64_t i0 = lo + t * tile + 2 * tid, i1 = i0 + 2 * T;
                double xv[4] = {xs[i0], xs[i0 + 1], xs[i1], xs[i1 + 1]}, vv[4] = {0.0, 0.0, 0.0, 0.0}, wv[4] = {0.0, 0.0, 0.0, 0.0};
                if (vs) { vv[0] = vs[i0]; vv[1] = vs[i0 + 1]; vv[2] = vs[i1]; vv[3] = vs[i1 + 1]; }
                if (WEIGHTED) { wv[0] = a.w[i0]; wv[1] = a.w[i0 + 1]; wv[2] = a.w[i1]; wv[3] = a.w[i1 + 1]; }
                deposit4(xv, vv, wv);
            }
            i = lo + nfull * tile + tid;
        }
        for (; i + 3 * T < hi; i += 4 * T) {
            double xv[4], vv[4], wv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                xv[u] = xs[i + u * T];
                wv[u] = WEIGHTED ? a.w[i + u * T] : 0.0;
                vv[u] = vs ? vs[i + u * T] : 0.0;
            }
            deposit4(xv, vv, wv);
        }
        for (; i < hi; i += T) {
            double xv = xs[i];
            const double wv = WEIGHTED ? a.w[i] : 0.0;
            if (vs) {
                double vv = vs[i];
                xv = __dadd_rn(xv, __dmul_rn(hdt, vv));
                double t = WEIGHTED ? wv * vv : vv;
                ksum = fma(t, vv, ksum);
                msum += t;
            }
            int c;
            double xi, b0, b1, b2, b3;
            locate(xv, g, c, xi);
            bspline3x6(xi, b0, b1, b2, b3);
            if (WEIGHTED) { b0 *= wv; b1 *= wv; b2 *= wv; b3 *= wv; }
            hist.rmw(c, b0, b1, b2, b3);
        }
        block_publish_hist<T>(histp, nh, a.part + ((size_t)s * a.nchunks + k) * nh);
        if (a.part_km) block_publish_km<T>(ksum, msum, red, a.part_km + ((size_t)s * a.nchunks + k) * 2);
        if (a.fuse_solve) {
            // saved-step update!: the last chunk of a sample to arrive sums the partials and solves the field here,
            // instead of a separate one-block-per-sample kernel behind this one
            __shared__ int s_last;
            if (tid == 0) {
                __threadfence();
                s_last = atomicAdd(a.tickets + s, 1) == a.nchunks - 1;
                if (s_last) a.tickets[s] = 0;
            }
            __syncthreads();
            if (s_last) {
                __threadfence();
                solve_sample<T>(a.sv, s, histp);   // the all-zero histogram doubles as scratch
                __syncthreads();
                for (int i = tid; i < solve_smem_doubles(nh, T); i += T) histp[i] = 0.0;
            }
            __syncthreads();
        }
    }
}

// Large-nh fallback: one RED.ADD.F64 per tap into part[s][:] (zeroed by the host); the optional
// K/M sums go to part_km[s][0][0..1] (all chunks zeroed by the host).
template <bool WEIGHTED>
__global__ void __launch_bounds__(256) k_deposit_global(const __grid_constant__ DepositArgs a)
{
    __shared__ double red[16];
    const Geom g = a.g;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int s = 0; s < a.nsamp; ++s) {
        const double *xs = a.x + (size_t)s * a.ldx;
        const double *vs = a.v ? a.v + (size_t)s * a.ldx : nullptr;
        const double hdt = a.v ? a.hdt[s] : 0.0;
        GlobalHist hist{a.part + (size_t)s * g.nh};
        double ksum = 0.0, msum = 0.0;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.np; i += stride) {
            double xv = xs[i];
            double wv = WEIGHTED ? a.w[i] : 0.0;
            if (vs) {
                double vv = vs[i];
                xv = __dadd_rn(xv, __dmul_rn(hdt, vv));
                double t = WEIGHTED ? wv * vv : vv;
                ksum = fma(t, vv, ksum);
                msum += t;
            }
            int c;
            double xi, b0, b1, b2, b3;
            locate(xv, g, c, xi);
            bsplinex6(xi, g.degree, b0, b1, b2, b3);
            if (WEIGHTED) { b0 *= wv; b1 *= wv; b2 *= wv; b3 *= wv; }
            hist.add4(c, g.nh, b0, b1, b2, b3);
        }
        if (a.part_km) {
            ksum = warp_sum(ksum);
            msum = warp_sum(msum);
            __syncthreads();
            if ((threadIdx.x & 31) == 0) {
                red[(threadIdx.x >> 5) * 2] = ksum;
                red[(threadIdx.x >> 5) * 2 + 1] = msum;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                double ks = 0.0, ms = 0.0;
                for (int wq = 0; wq < (int)(blockDim.x >> 5); ++wq) {
                    ks += red[wq * 2];
                    ms += red[wq * 2 + 1];
                }
                atomicAdd(a.part_km + (size_t)s * a.nchunks * 2, ks);
                atomicAdd(a.part_km + (size_t)s * a.nchunks * 2 + 1, ms);
            }
        }
    }
}

// Publish this rank's per-sample deposit sums to every rank's exchange buffer: data, system-scope fence, flags.
// One block of 1024 threads = G chunk-groups x n vector entries (n = nsamp*nh <= 1024); group g sums every
// G-th chunk partial with independent loads, groups are combined in a fixed order.
#ifndef RBM_NO_ENTRY_POINTS   // a second translation unit (recon.cu) includes this header for its device functions only
__global__ void __launch_bounds__(1024) k_reduce_publish(const double *part, int nsamp, int nchunks, int nh, PeerXch pub)
{
    __shared__ double grp[1024];
    const int n = nsamp * nh, tid = threadIdx.x;
    const unsigned long long pseq = xch_abs_seq(pub);
    const int pbuf = (int)(pseq & 1ull);
    if (n > 1024) {
        // many samples: few chunks per sample, one thread per entry walks them
        for (int idx = tid; idx < n; idx += 1024) {
            const int s = idx / nh, j = idx - s * nh;
            const double *p = part + (size_t)s * nchunks * nh + j;
            double acc = 0.0;
            for (int k = 0; k < nchunks; ++k) acc += p[(size_t)k * nh];
            for (int r = 0; r < pub.nranks; ++r) xch_data(pub.base[r], pbuf, pub.self)[idx] = acc;
        }
        __threadfence_system();
        __syncthreads();
        if (tid < pub.nranks) {
            volatile unsigned long long *flag = xch_flags(pub.base[tid]) + pbuf * XCH_R + pub.self;
            *flag = pseq;
            __threadfence_system();
        }
        return;
    }
    const int G = 1024 / n;
    const int g = tid / n, idx = tid - g * n;
    if (g < G) {
        const int s = idx / nh, j = idx - s * nh;
        const double *p = part + (size_t)s * nchunks * nh + j;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int k = g;
        for (; k + 3 * G < nchunks; k += 4 * G) {
            a0 += p[(size_t)k * nh];
            a1 += p[(size_t)(k + G) * nh];
            a2 += p[(size_t)(k + 2 * G) * nh];
            a3 += p[(size_t)(k + 3 * G) * nh];
        }
        for (; k < nchunks; k += G) a0 += p[(size_t)k * nh];
        grp[g * n + idx] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    if (tid < n) {
        double acc = 0.0;
        for (int q = 0; q < G; ++q) acc += grp[q * n + tid];
        for (int r = 0; r < pub.nranks; ++r) xch_data(pub.base[r], pbuf, pub.self)[tid] = acc;
    }
    __threadfence_system();
    __syncthreads();
    if (tid < pub.nranks) {
        volatile unsigned long long *flag = xch_flags(pub.base[tid]) + pbuf * XCH_R + pub.self;
        *flag = pseq;
        __threadfence_system();
    }
}
#endif

// end of a time loop: the next loop's exchanges are numbered from epoch + n (same on every rank)
#ifndef RBM_NO_ENTRY_POINTS   // a second translation unit (recon.cu) includes this header for its device functions only
__global__ void k_xch_advance(unsigned long long *epoch, unsigned long long n)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) *epoch += n;
}
#endif

// sum the per-chunk partials (before a cross-rank all-reduce): out[s][j] = sum_k part[s][k][j]
#ifndef RBM_NO_ENTRY_POINTS   // a second translation unit (recon.cu) includes this header for its device functions only
__global__ void k_reduce_part(const double *part, int nsamp, int nchunks, int width, double *out)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nsamp * width) return;
    int s = idx / width, j = idx - s * width;
    double acc = 0.0;
    const double *p = part + (size_t)s * nchunks * width + j;
    for (int k = 0; k < nchunks; ++k) acc += p[(size_t)k * width];
    out[idx] = acc;
}
#endif

// ------------------------------------------------------------------ gather / locate
struct GatherArgs {
    const double *x;     // [nsamp][ldx] (ldx = 0: shared)
    const double *v;     // optional, copied to V
    const double *coef;  // [nsamp][nh][3]
    double *A, *X, *V;   // slot of sample 0, any may be NULL
    int64_t snap_stride;
    int64_t np, ldx;
    int nsamp;
    Geom g;
};

#ifndef RBM_NO_ENTRY_POINTS   // a second translation unit (recon.cu) includes this header for its device functions only
__global__ void __launch_bounds__(256) k_gather(const __grid_constant__ GatherArgs a)
{
    const Geom g = a.g;
    const int s = blockIdx.y;
    const double *xs = a.x + (size_t)s * a.ldx;
    const double *vs = a.v ? a.v + (size_t)s * a.ldx : nullptr;
    const double *cf = a.coef + (size_t)s * g.nh * 3;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.np; i += stride) {
        double xv = xs[i];
        int c;
        double xi;
        locate(xv, g, c, xi);
        double al = __ldg(cf + c * 3), be = __ldg(cf + c * 3 + 1), ga = __ldg(cf + c * 3 + 2);
        double acc = fma(fma(ga, xi, be), xi, al);
        if (a.A) a.A[(size_t)s * a.snap_stride + i] = acc;
        if (a.X) a.X[(size_t)s * a.snap_stride + i] = xv;
        if (a.V && vs) a.V[(size_t)s * a.snap_stride + i] = vs[i];
    }
}
#endif

#ifndef RBM_NO_ENTRY_POINTS   // a second translation unit (recon.cu) includes this header for its device functions only
__global__ void __launch_bounds__(256) k_locate(const double *x, int64_t n, Geom g, int32_t *cell, double *xi_out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int c;
        double xi;
        locate(x[i], g, c, xi);
        if (cell) cell[i] = c;
        if (xi_out) xi_out[i] = xi;
    }
}
#endif

// state[s][0:np] = src[0:np] for every sample (the reference restarts every sample from
// the same `particles`: scripts/bump_on_tail_1_training.jl:87-97)
#ifndef RBM_NO_ENTRY_POINTS   // a second translation unit (recon.cu) includes this header for its device functions only
__global__ void __launch_bounds__(256) k_broadcast(const double *x0, const double *v0, double *x, double *v, int64_t np, int64_t ldp)
{
    const int s = blockIdx.y;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += stride) {
        x[(size_t)s * ldp + i] = x0[i];
        v[(size_t)s * ldp + i] = v0[i];
    }
}
#endif

}  // namespace rbm
