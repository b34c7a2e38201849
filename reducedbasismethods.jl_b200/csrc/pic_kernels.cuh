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
// summation order (bit-reproducible run to run).  Blocks publish per-chunk partial
// sums that k_solve adds in a fixed order.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace rbm {

struct Geom {
    double L, h, invL, invh;  // invh = RN(1/h), invL = RN(1/L)
    int nh;
    int exact_div;  // 1: use IEEE division for mod(x,L)/h (mantissa of h all ones)
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// floor for 0 <= s < 2^51 on the FP64 pipe (no F2I/I2F conversions):
// t = s + 2^52 rounds s to the nearest integer, whose value sits in the low mantissa bits.
__device__ __forceinline__ double floor_pos(double s, int &ifloor)
{
    const double M = 4503599627370496.0;  // 2^52
    double t = __dadd_rn(s, M);
    double r = __dadd_rn(t, -M);
    int c = __double2loint(t);
    if (r > s) {
        r = __dadd_rn(r, -1.0);
        c -= 1;
    }
    ifloor = c;
    return r;
}

// Julia mod(x, L) for L > 0, bit-exact with C fmod + sign fix-up:
// the remainder |x| - q*L is exactly representable, so one FMA with the right integer
// q = trunc(|x|/L) yields fmod exactly; q from a reciprocal multiply is off by at most
// one and is corrected by the sign/size of the FMA result.
__device__ __forceinline__ double julia_mod(double x, const Geom &g)
{
    double ax = fabs(x);
    double t = __dmul_rn(ax, g.invL);
    double r;
    if (t < 1.0e15) {
        int qi;
        double q = floor_pos(t, qi);
        r = __fma_rn(-q, g.L, ax);
        if (r < 0.0) {
            q = __dadd_rn(q, -1.0);
            r = __fma_rn(-q, g.L, ax);
        } else if (r >= g.L) {
            q = __dadd_rn(q, 1.0);
            r = __fma_rn(-q, g.L, ax);
        }
        if (x < 0.0 && r != 0.0) r = __dadd_rn(g.L, -r);  // (-r) + L
    } else {  // huge or non-finite argument: library fmod (exact)
        r = fmod(x, g.L);
        if (r < 0.0) r = __dadd_rn(r, g.L);
        if (r == 0.0) r = 0.0;
    }
    return r;
}

// s = mod(x,L)/h ; c = floor(s) clamped to [0, nh-1] ; xi = s - c.
// The division is Markstein's correction of r*RN(1/h) (correctly rounded, 3 FP64 ops);
// tests/test_oracle_pic.py and tests/test_gpu_pic.py check it against IEEE division.
__device__ __forceinline__ void locate(double x, const Geom &g, int &c, double &xi)
{
    double r = julia_mod(x, g);
    double s;
    if (g.exact_div) {
        s = __ddiv_rn(r, g.h);
    } else {
        double q0 = __dmul_rn(r, g.invh);
        double e = __fma_rn(-g.h, q0, r);
        s = __fma_rn(e, g.invh, q0);
    }
    int ci;
    double fl = floor_pos(s, ci);
    if (ci > g.nh - 1) {
        ci = g.nh - 1;
        fl = (double)(g.nh - 1);
    }
    if (ci < 0) {  // only for NaN input; keeps the shared-memory index in range
        ci = 0;
        fl = 0.0;
    }
    c = ci;
    xi = __dadd_rn(s, -fl);
}

// 6*B_m(xi), m = 0..3 (the 1/6 and a uniform weight are applied once in k_solve)
__device__ __forceinline__ void bspline3x6(double xi, double &b0, double &b1, double &b2, double &b3)
{
    double u = 1.0 - xi;
    double x2 = xi * xi;
    b0 = u * u * u;
    b3 = x2 * xi;
    b1 = fma(x2, fma(3.0, xi, -6.0), 4.0);
    b2 = fma(fma(fma(-3.0, xi, 3.0), xi, 3.0), xi, 1.0);
}

// ---------------------------------------------------------------------------------
// Thread-private histogram in shared memory: hist[bin*T + tid].
struct SmemHist {
    double *h;
    int T, tid;
    __device__ __forceinline__ void add4(int c, int nh, double b0, double b1, double b2, double b3)
    {
        int j0 = c, j1 = c + 1, j2 = c + 2, j3 = c + 3;
        if (j1 >= nh) j1 -= nh;
        if (j2 >= nh) j2 -= nh;
        if (j3 >= nh) j3 -= nh;
        double *p0 = h + j0 * T + tid, *p1 = h + j1 * T + tid, *p2 = h + j2 * T + tid,
               *p3 = h + j3 * T + tid;
        double h0 = *p0, h1 = *p1, h2 = *p2, h3 = *p3;  // four distinct bins (nh >= 4)
        *p0 = h0 + b0;
        *p1 = h1 + b1;
        *p2 = h2 + b2;
        *p3 = h3 + b3;
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
    int rep;       // replication of the gather table across banks (power of two <= 16)
    int snap_vec;  // 1: snapshot slots are 16-byte aligned for every sample
    int do_mid;    // 0: skip the deposit (last step)
    Geom g;
};

template <bool WEIGHTED, class Hist>
__device__ __forceinline__ void deposit_one(double xm, double w, const Geom &g, Hist &hist)
{
    int c;
    double xi;
    locate(xm, g, c, xi);
    double b0, b1, b2, b3;
    bspline3x6(xi, b0, b1, b2, b3);
    if (WEIGHTED) {
        b0 *= w;
        b1 *= w;
        b2 *= w;
        b3 *= w;
    }
    hist.add4(c, g.nh, b0, b1, b2, b3);
}

// One particle: drift, gather, kick, drift (bit-exact products/sums, never contracted),
// then the next step's half drift + deposit.
template <bool SAVE, bool WEIGHTED, class Hist>
__device__ __forceinline__ void push_one(double &x, double &v, double w, double hdt, double dte,
                                         const Geom &g, const double *tab, int rep, int lane_r,
                                         Hist &hist, bool do_mid, double &acc_out, double &ksum,
                                         double &msum)
{
    double xp = __dadd_rn(x, __dmul_rn(hdt, v));
    int c;
    double xi;
    locate(xp, g, c, xi);
    const double *tp = tab + (c * 3) * rep + lane_r;
    double al = tp[0], be = tp[rep], ga = tp[2 * rep];
    double acc = fma(fma(ga, xi, be), xi, al);
    double vn = __dadd_rn(v, __dmul_rn(dte, acc));
    double d = __dmul_rn(hdt, vn);
    double xn = __dadd_rn(xp, d);
    if (do_mid) deposit_one<WEIGHTED>(__dadd_rn(xn, d), w, g, hist);
    if (SAVE) {
        double wv = WEIGHTED ? w * vn : vn;
        ksum = fma(wv, vn, ksum);
        msum += wv;
    }
    x = xn;
    v = vn;
    acc_out = acc;
}

// shared-memory layout of k_step / k_deposit: hist[nh*T] | tab[nh*3*rep] | red[64]
__device__ __forceinline__ void block_publish_hist(double *hist, int nh, int T, double *dst)
{
    // dst[bin] = sum_t hist[bin][t] in a fixed order; hist is zeroed for the next item
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = T >> 5;
    for (int bin = warp; bin < nh; bin += nwarps) {
        double s = 0.0;
        double *row = hist + bin * T;
        for (int t = lane; t < T; t += 32) {
            s += row[t];
            row[t] = 0.0;
        }
        s = warp_sum(s);
        if (lane == 0) dst[bin] = s;
    }
    __syncthreads();
}

// GHIST = false: thread-private shared-memory histograms, partials part[s][chunk][:].
// GHIST = true : large-nh fallback, RED.ADD.F64 into part[s][:] (zeroed by the host).
template <bool SAVE, bool WEIGHTED, bool GHIST>
__global__ void __launch_bounds__(512, 1) k_step(const __grid_constant__ StepArgs a)
{
    extern __shared__ double smem[];
    const int T = blockDim.x, tid = threadIdx.x, nh = a.g.nh, rep = a.rep;
    double *histp = smem;
    double *tab = smem + (GHIST ? 0 : (size_t)nh * T);
    double *red = tab + nh * 3 * rep;
    const int lane_r = tid & (rep - 1);
    if (!GHIST)
        for (int i = tid; i < nh * T; i += T) histp[i] = 0.0;
    typename std::conditional<GHIST, GlobalHist, SmemHist>::type hist;
    if constexpr (!GHIST) {
        hist.h = histp;
        hist.T = T;
        hist.tid = tid;
    }
    const Geom g = a.g;
    const int items = a.nsamp * a.nchunks;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int s = item / a.nchunks, k = item - s * a.nchunks;
        __syncthreads();
        for (int i = tid; i < nh * 3 * rep; i += T) tab[i] = a.coef[(size_t)s * nh * 3 + i / rep];
        __syncthreads();
        const double hdt = a.hdt[s], dte = a.dte[s];
        const int64_t lo = (int64_t)k * a.chunk_len;
        int64_t hi = lo + a.chunk_len;
        if (hi > a.np) hi = a.np;
        double *xs = a.x + (size_t)s * a.ldp, *vs = a.v + (size_t)s * a.ldp;
        double *Xs = a.X ? a.X + (size_t)s * a.snap_stride : nullptr;
        double *Vs = a.V ? a.V + (size_t)s * a.snap_stride : nullptr;
        double *As = a.A ? a.A + (size_t)s * a.snap_stride : nullptr;
        if constexpr (GHIST) hist.h = a.part + (size_t)s * nh;
        double ksum = 0.0, msum = 0.0;
        const bool do_mid = a.do_mid != 0;
        const int64_t tile = 4 * (int64_t)T;
        const int64_t nfull = hi > lo ? (hi - lo) / tile : 0;
        // ---- full tiles: two double2 per array per thread, next tile prefetched ----
        double2 cx0, cx1, cv0, cv1, cw0, cw1;
        cw0 = cw1 = make_double2(0.0, 0.0);
        if (nfull > 0) {
            const int64_t i0 = lo + 2 * tid, i1 = i0 + 2 * T;
            cx0 = *reinterpret_cast<const double2 *>(xs + i0);
            cx1 = *reinterpret_cast<const double2 *>(xs + i1);
            cv0 = *reinterpret_cast<const double2 *>(vs + i0);
            cv1 = *reinterpret_cast<const double2 *>(vs + i1);
            if (WEIGHTED) {
                cw0 = *reinterpret_cast<const double2 *>(a.w + i0);
                cw1 = *reinterpret_cast<const double2 *>(a.w + i1);
            }
        }
        for (int64_t t = 0; t < nfull; ++t) {
            const int64_t i0 = lo + t * tile + 2 * tid, i1 = i0 + 2 * T;
            double2 nx0, nx1, nv0, nv1, nw0, nw1;
            nw0 = nw1 = make_double2(0.0, 0.0);
            if (t + 1 < nfull) {
                nx0 = *reinterpret_cast<const double2 *>(xs + i0 + tile);
                nx1 = *reinterpret_cast<const double2 *>(xs + i1 + tile);
                nv0 = *reinterpret_cast<const double2 *>(vs + i0 + tile);
                nv1 = *reinterpret_cast<const double2 *>(vs + i1 + tile);
                if (WEIGHTED) {
                    nw0 = *reinterpret_cast<const double2 *>(a.w + i0 + tile);
                    nw1 = *reinterpret_cast<const double2 *>(a.w + i1 + tile);
                }
            }
            double2 a0, a1;
            push_one<SAVE, WEIGHTED>(cx0.x, cv0.x, cw0.x, hdt, dte, g, tab, rep, lane_r, hist, do_mid, a0.x, ksum, msum);
            push_one<SAVE, WEIGHTED>(cx0.y, cv0.y, cw0.y, hdt, dte, g, tab, rep, lane_r, hist, do_mid, a0.y, ksum, msum);
            push_one<SAVE, WEIGHTED>(cx1.x, cv1.x, cw1.x, hdt, dte, g, tab, rep, lane_r, hist, do_mid, a1.x, ksum, msum);
            push_one<SAVE, WEIGHTED>(cx1.y, cv1.y, cw1.y, hdt, dte, g, tab, rep, lane_r, hist, do_mid, a1.y, ksum, msum);
            *reinterpret_cast<double2 *>(xs + i0) = cx0;
            *reinterpret_cast<double2 *>(xs + i1) = cx1;
            *reinterpret_cast<double2 *>(vs + i0) = cv0;
            *reinterpret_cast<double2 *>(vs + i1) = cv1;
            if (SAVE) {
                if (a.snap_vec) {
                    if (Xs) { __stcs(reinterpret_cast<double2 *>(Xs + i0), cx0); __stcs(reinterpret_cast<double2 *>(Xs + i1), cx1); }
                    if (Vs) { __stcs(reinterpret_cast<double2 *>(Vs + i0), cv0); __stcs(reinterpret_cast<double2 *>(Vs + i1), cv1); }
                    if (As) { __stcs(reinterpret_cast<double2 *>(As + i0), a0); __stcs(reinterpret_cast<double2 *>(As + i1), a1); }
                } else {
                    if (Xs) { Xs[i0] = cx0.x; Xs[i0 + 1] = cx0.y; Xs[i1] = cx1.x; Xs[i1 + 1] = cx1.y; }
                    if (Vs) { Vs[i0] = cv0.x; Vs[i0 + 1] = cv0.y; Vs[i1] = cv1.x; Vs[i1 + 1] = cv1.y; }
                    if (As) { As[i0] = a0.x; As[i0 + 1] = a0.y; As[i1] = a1.x; As[i1 + 1] = a1.y; }
                }
            }
            cx0 = nx0; cx1 = nx1; cv0 = nv0; cv1 = nv1;
            if (WEIGHTED) { cw0 = nw0; cw1 = nw1; }
        }
        // ---- remainder (< 4T particles), scalar ----
        for (int64_t i = lo + nfull * tile + tid; i < hi; i += T) {
            double xv = xs[i], vv = vs[i], acc;
            double wv = WEIGHTED ? a.w[i] : 0.0;
            push_one<SAVE, WEIGHTED>(xv, vv, wv, hdt, dte, g, tab, rep, lane_r, hist, do_mid, acc, ksum, msum);
            xs[i] = xv;
            vs[i] = vv;
            if (SAVE) {
                if (Xs) Xs[i] = xv;
                if (Vs) Vs[i] = vv;
                if (As) As[i] = acc;
            }
        }
        if (!GHIST && do_mid) block_publish_hist(histp, nh, T, a.part + ((size_t)s * a.nchunks + k) * nh);
        if (SAVE) {
            ksum = warp_sum(ksum);
            msum = warp_sum(msum);
            if ((tid & 31) == 0) {
                red[(tid >> 5) * 2] = ksum;
                red[(tid >> 5) * 2 + 1] = msum;
            }
            __syncthreads();
            if (tid == 0) {
                double ks = 0.0, ms = 0.0;
                for (int wq = 0; wq < (T >> 5); ++wq) {
                    ks += red[wq * 2];
                    ms += red[wq * 2 + 1];
                }
                a.part_km[((size_t)s * a.nchunks + k) * 2] = ks;
                a.part_km[((size_t)s * a.nchunks + k) * 2 + 1] = ms;
            }
        }
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
    Geom g;
};

template <bool WEIGHTED>
__global__ void __launch_bounds__(512, 1) k_deposit(const __grid_constant__ DepositArgs a)
{
    extern __shared__ double smem[];
    const int T = blockDim.x, tid = threadIdx.x, nh = a.g.nh;
    double *histp = smem;
    double *red = smem + (size_t)nh * T;
    for (int i = tid; i < nh * T; i += T) histp[i] = 0.0;
    SmemHist hist{histp, T, tid};
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
#pragma unroll 4
        for (int64_t i = lo + tid; i < hi; i += T) {
            double xv = xs[i];
            double wv = WEIGHTED ? a.w[i] : 0.0;
            if (vs) {
                double vv = vs[i];
                xv = __dadd_rn(xv, __dmul_rn(hdt, vv));
                if (a.part_km) {
                    double t = WEIGHTED ? wv * vv : vv;
                    ksum = fma(t, vv, ksum);
                    msum += t;
                }
            }
            deposit_one<WEIGHTED>(xv, wv, g, hist);
        }
        block_publish_hist(histp, nh, T, a.part + ((size_t)s * a.nchunks + k) * nh);
        if (a.part_km) {
            ksum = warp_sum(ksum);
            msum = warp_sum(msum);
            if ((tid & 31) == 0) {
                red[(tid >> 5) * 2] = ksum;
                red[(tid >> 5) * 2 + 1] = msum;
            }
            __syncthreads();
            if (tid == 0) {
                double ks = 0.0, ms = 0.0;
                for (int wq = 0; wq < (T >> 5); ++wq) {
                    ks += red[wq * 2];
                    ms += red[wq * 2 + 1];
                }
                a.part_km[((size_t)s * a.nchunks + k) * 2] = ks;
                a.part_km[((size_t)s * a.nchunks + k) * 2 + 1] = ms;
            }
            __syncthreads();
        }
    }
}

// Large-nh fallback: one RED.ADD.F64 per tap into part[s][0][:] (nchunks == 1, zeroed by host)
template <bool WEIGHTED>
__global__ void __launch_bounds__(256) k_deposit_global(const __grid_constant__ DepositArgs a)
{
    const Geom g = a.g;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int s = 0; s < a.nsamp; ++s) {
        const double *xs = a.x + (size_t)s * a.ldx;
        const double *vs = a.v ? a.v + (size_t)s * a.ldx : nullptr;
        const double hdt = a.v ? a.hdt[s] : 0.0;
        GlobalHist hist{a.part + (size_t)s * g.nh};
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.np; i += stride) {
            double xv = xs[i];
            if (vs) xv = __dadd_rn(xv, __dmul_rn(hdt, vs[i]));
            deposit_one<WEIGHTED>(xv, WEIGHTED ? a.w[i] : 0.0, g, hist);
        }
    }
}

// ------------------------------------------------------------------ Poisson solve
struct SolveArgs {
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
};

// One block per sample: rho = rhs - mean; phi0 = -pinv(K_s) rho; phi = phi0/chi^2;
// per-cell quadratic of phi'(x); W = chi^2/2 phi'K_s phi.
__global__ void __launch_bounds__(256) k_solve(const __grid_constant__ SolveArgs a)
{
    extern __shared__ double sm[];
    const int nh = a.nh, T = blockDim.x, tid = threadIdx.x, s = blockIdx.x;
    double *rho = sm, *phi = sm + nh, *red = sm + 2 * nh;  // red[64]
    const double chi = a.chi[s];
    for (int j = tid; j < nh; j += T) {
        double acc = 0.0;
        const double *p = a.part + (size_t)s * a.nchunks * nh + j;
        for (int k = 0; k < a.nchunks; ++k) acc += p[(size_t)k * nh];
        acc *= a.scale;
        rho[j] = acc;
        if (a.rhs_out) a.rhs_out[(size_t)s * nh + j] = acc;
    }
    __syncthreads();
    // mean (fixed order: strided partials, then serial over the block's lanes)
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
    for (int j = tid; j < nh; j += T) {
        double acc = 0.0;
        for (int i = 0; i < nh; ++i) {
            int d = j - i;
            if (d < 0) d += nh;
            acc = fma(a.pinv[d], rho[i], acc);
        }
        double ph = -acc * ichi2;
        phi[j] = ph;
        if (a.phi_out) a.phi_out[(size_t)s * a.phi_stride + j] = ph;
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
            o[0] = 0.5 * (p2 - p0) * ih;
            o[1] = ((p0 - 2.0 * p1) + p2) * ih;
            o[2] = 0.5 * ((p3 - p0) + 3.0 * (p1 - p2)) * ih;
        }
    }
    if (a.W_out) {
        const double band[7] = {-1.0, -24.0, -15.0, 80.0, -15.0, -24.0, -1.0};
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
            a.W_out[(size_t)s * a.w_stride] = chi * chi * 0.5 * tot / (120.0 * a.h);
            if (a.part_km && a.K_out && a.M_out) {
                double ks = 0.0, ms = 0.0;
                const double *q = a.part_km + (size_t)s * a.km_chunks * 2;
                for (int k = 0; k < a.km_chunks; ++k) {
                    ks += q[2 * k];
                    ms += q[2 * k + 1];
                }
                a.K_out[(size_t)s * a.w_stride] = 0.5 * a.km_scale * ks;
                a.M_out[(size_t)s * a.w_stride] = a.km_scale * ms;
            }
        }
    }
}

// sum the per-chunk partials (before a cross-rank all-reduce): out[s][j] = sum_k part[s][k][j]
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

// state[s][0:np] = src[0:np] for every sample (the reference restarts every sample from
// the same `particles`: scripts/bump_on_tail_1_training.jl:87-97)
__global__ void __launch_bounds__(256) k_broadcast(const double *x0, const double *v0, double *x, double *v, int64_t np, int64_t ldp)
{
    const int s = blockIdx.y;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += stride) {
        x[(size_t)s * ldp + i] = x0[i];
        v[(size_t)s * ldp + i] = v0[i];
    }
}

}  // namespace rbm
