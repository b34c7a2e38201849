"""oracle/pic_numpy.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy restatement of the full-order 1D1V Vlasov-Poisson PIC path of
JuliaRCM/ReducedBasisMethods.jl, written independently of oracle/pic_oracle.c so
that the two can be checked against each other (tests/test_oracle_pic.py).

PARITY UNPINNED: the arithmetic lives in VlasovMethods.jl / PoissonSolvers.jl /
ParticleMethods.jl, which are registry dependencies absent from /root/reference
(Project.toml:19,21,26,36,38,41; no Manifest.toml) and no reference test pins it
(test/trainingset_tests.jl:21-22 only constructs the types).  The restatement
follows the reference's call sites and in-tree twin (file:line in each docstring)
and SURVEY.md Appendix A.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.
"""
from __future__ import annotations

import numpy as np

class Conventions:
    """The conventions SURVEY.md App. A.4 lists as NOT pinned by anything in /root/reference, as switches.  The defaults
    are what the CUDA path implements.  tests/test_reference_vectors.py determines the real packages' values from
    tests/golden/ref_julia_outputs.bin (tools/reference_vectors.jl) when that file exists.

      phi_sign    +1: stored coefficients are those of phi with a = +d(phi)/dx;  -1: of -phi      (A.4 (1))
      gauge       constant added to the zero-mean coefficients                                    (A.4 (2))
      tap_shift   cyclic shift of the stored coefficient vector (which cell tap 0 belongs to)     (A.4 (3))
      a_at_save   "kick": IC.A at a save is the acceleration used in that step's kick (initial save: field at x0);
                  "post_update": the field of the update! at the saved positions, gathered there   (A.4 (4))
    A, W, K, M, X, V are invariant under the first three; only the stored Phi changes."""

    def __init__(self, phi_sign=1.0, gauge=0.0, tap_shift=0, a_at_save="kick"):
        assert a_at_save in ("kick", "post_update")
        self.phi_sign, self.gauge, self.tap_shift, self.a_at_save = float(phi_sign), float(gauge), int(tap_shift), a_at_save

    def stored_phi(self, phi):
        """canonical coefficients (zero mean, tap 0 = the particle's own cell, a = +phi') -> as stored under these conventions"""
        return self.phi_sign * np.roll(phi, self.tap_shift, axis=-1) + self.gauge

    def __repr__(self):
        return (f"Conventions(phi_sign={self.phi_sign:+.0f}, gauge={self.gauge:.3e}, tap_shift={self.tap_shift}, "
                f"a_at_save={self.a_at_save!r})")


DEFAULT = Conventions()

KS_BAND = np.array([-1.0, -24.0, -15.0, 80.0, -15.0, -24.0, -1.0]) / 120.0
# stiffness bands of periodic B-splines of degree 1 and 2 (int B_i' B_j' dx * h), padded to seven entries
KS_BANDS = {1: np.array([0.0, 0.0, -1.0, 2.0, -1.0, 0.0, 0.0]),
            2: np.array([0.0, -1.0, -2.0, 6.0, -2.0, -1.0, 0.0]) / 6.0,
            3: KS_BAND}
MS_BAND = np.array([1.0, 120.0, 1191.0, 2416.0, 1191.0, 120.0, 1.0]) / 5040.0


def julia_mod(x, L):
    """Julia ``mod(x::Float64, L)``: fmod + sign fix-up (periodicity lives in the
    evaluation, x is never wrapped: src/particles/electric_field.jl:23-24)."""
    r = np.fmod(x, L)
    neg = r < 0.0
    r = np.where(neg, r + L, r)
    return np.where(r == 0.0, 0.0, r)


def locate(x, L, nh):
    """s = mod(x,L)/h, c = floor(s) clamped to nh-1, xi = s - c."""
    h = L / nh
    s = julia_mod(np.asarray(x, dtype=np.float64), L) / h
    c = np.floor(s).astype(np.int32)
    c = np.minimum(c, nh - 1)
    xi = s - c.astype(np.float64)
    return c, xi


def bspline(xi, degree=3):
    """Values of the degree+1 non-zero periodic B-splines at offset xi in a cell, taps m = 0..degree on the cells
    c..c+degree (tap 0 = the particle's own cell), padded with zeros to four taps."""
    if degree == 3:
        return bspline3(xi)
    u = 1.0 - xi
    z = np.zeros_like(xi)
    if degree == 2:
        return np.stack([u * u / 2.0, (-2.0 * xi * xi + 2.0 * xi + 1.0) / 2.0, xi * xi / 2.0, z])
    if degree == 1:
        return np.stack([u, xi, z, z])
    raise ValueError("degree must be 1, 2 or 3")


def bspline_deriv(xi, h, degree=3):
    if degree == 3:
        return bspline3_deriv(xi, h)
    z = np.zeros_like(xi)
    if degree == 2:
        return np.stack([-(1.0 - xi) / h, (1.0 - 2.0 * xi) / h, xi / h, z])
    if degree == 1:
        return np.stack([-np.ones_like(xi) / h, np.ones_like(xi) / h, z, z])
    raise ValueError("degree must be 1, 2 or 3")


def bspline3(xi):
    u = 1.0 - xi
    return np.stack([
        u * u * u / 6.0,
        (3.0 * xi * xi * xi - 6.0 * xi * xi + 4.0) / 6.0,
        (-3.0 * xi * xi * xi + 3.0 * xi * xi + 3.0 * xi + 1.0) / 6.0,
        xi * xi * xi / 6.0,
    ])


def bspline3_deriv(xi, h):
    u = 1.0 - xi
    return np.stack([
        -(u * u) / 2.0 / h,
        (9.0 * xi * xi - 12.0 * xi) / 6.0 / h,
        (-9.0 * xi * xi + 6.0 * xi + 3.0) / 6.0 / h,
        (xi * xi) / 2.0 / h,
    ])


def deposit(x, w, L, nh, degree=3):
    """rhs_j = sum_p w_p B_j(x_p)  (PoissonSolvers ``rhs_particles_PBSBasis``,
    imported at src/particles/time_marching.jl:3).  Accumulated in longdouble."""
    c, xi = locate(x, L, nh)
    b = bspline(xi, degree)
    w = np.broadcast_to(np.asarray(w, dtype=np.float64), c.shape)
    rhs = np.zeros(nh, dtype=np.longdouble)
    for m in range(4):
        np.add.at(rhs, (c + m) % nh, (w * b[m]).astype(np.longdouble))
    return rhs.astype(np.float64)


def circulant(band, nh, scale):
    K = np.zeros((nh, nh))
    for i in range(nh):
        for d in range(-3, 4):
            K[i, (i + d) % nh] += band[d + 3] * scale
    return K


def stiffness(nh, L, degree=3):
    """K_s = (1/h) circ(-1,-24,-15,80,-15,-24,-1)/120  (SURVEY.md A.2); degree 2: (1/h) circ(-1,-2,6,-2,-1)/6; 1: (1/h) circ(-1,2,-1)."""
    return circulant(KS_BANDS[degree], nh, nh / L)


def mass(nh, L):
    """M_s = h circ(1,120,1191,2416,1191,120,1)/5040  (SURVEY.md A.2)."""
    return circulant(MS_BAND, nh, L / nh)


def poisson(rhs, nh, L, chi, degree=3):
    """`update!` after the deposit: remove the mean, solve K_s phi0 = -(rhs-mean)
    in the zero-mean gauge ((K+R) form of src/gridbased/poisson.jl:34-43),
    phi = phi0/chi^2, W = chi^2 * 1/2 phi'K_s phi
    (notebooks/VP-ROM-Plotting.ipynb cell 24)."""
    K = stiffness(nh, L, degree)
    rho = rhs - rhs.mean()
    phi0 = np.linalg.solve(K + 1.0 / nh, -rho)
    phi = phi0 / (chi * chi)
    W = chi * chi * 0.5 * phi @ (K @ phi)
    return phi, W


def gather(phi, x, L, nh, degree=3):
    """a_p = sum_j phi_j B_j'(x_p)  (``eval_deriv_PBSBasis``; `efield!` at
    src/particles/electric_field.jl:27-31)."""
    c, xi = locate(x, L, nh)
    d = bspline_deriv(xi, L / nh, degree)
    a = np.zeros_like(xi)
    for m in range(4):
        a = a + phi[(c + m) % nh] * d[m]
    return a


def integrate(x0, v0, w, L, nh, chi, dt, nt, ns, conv: Conventions = DEFAULT, degree=3):
    """integrate_vp! for one sample: src/particles/time_marching.jl:119-149 with
    Psi = I; save_solution at :81-105; call site scripts/bump_on_tail_1_training.jl:97."""
    np_ = x0.shape[0]
    nsave = nt // (ns - 1)
    x = np.array(x0, dtype=np.float64)
    v = np.array(v0, dtype=np.float64)
    w = np.broadcast_to(np.asarray(w, dtype=np.float64), x.shape)
    X = np.zeros((ns, np_)); V = np.zeros((ns, np_)); A = np.zeros((ns, np_))
    Phi = np.zeros((ns, nh)); W = np.zeros(ns); K = np.zeros(ns); M = np.zeros(ns)
    dte = dt * chi
    hdt = 0.5 * dte
    a = np.zeros_like(x)
    ts = 0
    for it in range(nt + 1):
        if it > 0:
            x = x + hdt * v
            phi, _ = poisson(deposit(x, w, L, nh, degree), nh, L, chi, degree)
            a = gather(phi, x, L, nh, degree)
            v = v + dte * a
            x = x + hdt * v
        if it % nsave == 0 and ts < ns:
            phi, Wt = poisson(deposit(x, w, L, nh, degree), nh, L, chi, degree)
            if it == 0 or conv.a_at_save == "post_update":
                a_save = gather(phi, x, L, nh, degree)
                if it == 0:
                    a = a_save
            else:
                a_save = a
            X[ts] = x; V[ts] = v; A[ts] = a_save; Phi[ts] = conv.stored_phi(phi)
            W[ts] = Wt
            K[ts] = float(np.sum((v * w * v).astype(np.longdouble)) / 2)
            M[ts] = float(np.sum((w * v).astype(np.longdouble)))
            ts += 1
    return dict(X=X, V=V, A=A, Phi=Phi, W=W, K=K, M=M)
