"""PoissonSolverPBSplines / ScaledPoissonField -- the spline Poisson field of PoissonSolvers.jl /
VlasovMethods.jl as the reference uses it, backed by librbm_b200.

Reference surface: ctor ``PoissonSolverPBSplines(p, nh, L)`` (scripts/bump_on_tail_1_training.jl:73),
fields ``p``, ``L`` (src/particles/poisson.jl:7-8), ``length`` = nh (src/trainingset.jl:21), ``==``
(test/trainingset_tests.jl:53); ``ScaledPoissonField(poisson, chi)`` (scripts/...1_training.jl:94)
and the ElectricField protocol update!/efield!/energy/coefficients plus the callable form
(src/particles/electric_field.jl:22-35, src/particles/time_marching.jl:95-99,140)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .engine import default_context


class PoissonSolverPBSplines:
    """Parameter carrier (p, nh, L); the solve itself lives in the CUDA library."""

    def __init__(self, p: int, nh: int, L: float):
        self.p = int(p)
        self.nh = int(nh)
        self.L = float(L)

    def __len__(self):
        return self.nh

    def __eq__(self, o):
        return isinstance(o, PoissonSolverPBSplines) and (self.p, self.nh, self.L) == (o.p, o.nh, o.L)


class PICHandle:
    """rbm_pic: device-side particles + spline field for one (np, nh, p, L)."""

    def __init__(self, np_: int, poisson: PoissonSolverPBSplines, ctx=None):
        self.ctx = ctx or default_context()
        self.lib = self.ctx.lib
        self.np = int(np_)
        self.poisson = poisson
        h = C.c_void_p()
        self.ctx.check(self.lib.rbm_pic_create(self.ctx.h, self.np, poisson.nh, poisson.p, poisson.L, C.byref(h)))
        self.h = h
        self._keep = []

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.lib.rbm_pic_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_particles(self, x, v, w):
        """w: array of length np, or a float (uniform weight)."""
        wa, wu = (None, float(w)) if np.isscalar(w) else (w, 0.0)
        for name, a in (("x", x), ("v", v), ("w", wa)):
            _lib.check_buffer(a, self.np, name, writable=False)
        self.ctx.check(self.lib.rbm_pic_set_particles(self.h, _lib.ptr(x), _lib.ptr(v), _lib.ptr(wa), wu))

    def set_conventions(self, phi_sign=1.0, tap_shift=0, gauge=0.0, a_at_save="kick"):
        """The output conventions nothing in the reference pins (rbm_pic_set_conventions): sign, cyclic tap shift and
        gauge of the stored Phi; A at a save = the kick's acceleration or the post-update! field."""
        if a_at_save not in ("kick", "post_update"):
            raise ValueError("a_at_save must be 'kick' or 'post_update'")
        self.ctx.check(self.lib.rbm_pic_set_conventions(self.h, float(phi_sign), int(tap_shift), float(gauge),
                                                        1 if a_at_save == "post_update" else 0))

    def run(self, chi, dt, nt, ns, X=None, V=None, A=None, Phi=None, W=None, K=None, M=None):
        chi = np.ascontiguousarray(chi, dtype=np.float64).reshape(-1)
        nparam, nh = chi.shape[0], self.poisson.nh
        for name, a in (("X", X), ("V", V), ("A", A)):
            _lib.check_buffer(a, nparam * int(ns) * self.np, name)
        _lib.check_buffer(Phi, nparam * int(ns) * nh, "Phi")
        for name, a in (("W", W), ("K", K), ("M", M)):
            _lib.check_buffer(a, nparam * int(ns), name)
        self.ctx.check(self.lib.rbm_pic_run(self.h, chi.ctypes.data, chi.shape[0], float(dt), int(nt), int(ns),
                                            *[_lib.ptr(a) for a in (X, V, A, Phi, W, K, M)]))

    def locate(self, x, n=None):
        n = int(n if n is not None else x.shape[0])
        cell = np.empty(n, dtype=np.int32)
        xi = np.empty(n, dtype=np.float64)
        self.ctx.check(self.lib.rbm_locate(self.h, _lib.ptr(x), n, cell.ctypes.data, xi.ctypes.data))
        return cell, xi

    def deposit(self, x, w, n=None):
        n = int(n if n is not None else x.shape[0])
        wa, wu = (None, float(w)) if np.isscalar(w) else (w, 0.0)
        rhs = np.empty(self.poisson.nh, dtype=np.float64)
        self.ctx.check(self.lib.rbm_deposit(self.h, _lib.ptr(x), _lib.ptr(wa), wu, n, rhs.ctypes.data))
        return rhs

    def step(self, x, v, w, chi, dt):
        """One drift-kick-drift step on copies of (x, v); returns (x1, v1, rhs, phi, a)."""
        x = np.array(x, dtype=np.float64); v = np.array(v, dtype=np.float64)
        n = x.shape[0]
        wa, wu = (None, float(w)) if np.isscalar(w) else (np.ascontiguousarray(w, dtype=np.float64), 0.0)
        rhs = np.empty(self.poisson.nh); phi = np.empty(self.poisson.nh); a = np.empty(n)
        self.ctx.check(self.lib.rbm_pic_step(self.h, x.ctypes.data, v.ctypes.data, _lib.ptr(wa), wu, n,
                                             float(chi), float(dt), rhs.ctypes.data, phi.ctypes.data, a.ctypes.data))
        return x, v, rhs, phi, a

    # ElectricField protocol ---------------------------------------------------------
    def field_update(self, x, w, chi, n=None):
        n = int(n if n is not None else x.shape[0])
        wa, wu = (None, float(w)) if np.isscalar(w) else (w, 0.0)
        self.ctx.check(self.lib.rbm_field_update(self.h, _lib.ptr(x), _lib.ptr(wa), wu, n, float(chi)))

    def field_eval(self, x, E=None, n=None):
        n = int(n if n is not None else x.shape[0])
        if E is None:
            E = np.empty(n, dtype=np.float64)
        self.ctx.check(self.lib.rbm_field_eval(self.h, _lib.ptr(x), n, _lib.ptr(E)))
        return E

    def field_energy(self):
        W = C.c_double()
        self.ctx.check(self.lib.rbm_field_energy(self.h, C.byref(W)))
        return W.value

    def field_coefficients(self):
        phi = np.empty(self.poisson.nh, dtype=np.float64)
        self.ctx.check(self.lib.rbm_field_coefficients(self.h, phi.ctypes.data))
        return phi


class ScaledPoissonField:
    """``ScaledPoissonField(poisson, chi)`` backed by the GPU (the Julia side names it
    B200PoissonField <: ElectricField).  update!/efield!/energy/coefficients and f(E, x, w, t)."""

    def __init__(self, poisson: PoissonSolverPBSplines, chi: float, ctx=None, handle: PICHandle | None = None):
        self.poisson = poisson
        self.chi = float(chi)
        self.handle = handle or PICHandle(1, poisson, ctx)

    def update(self, x, w, t=0.0):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        wv = w if np.isscalar(w) else np.ascontiguousarray(w, dtype=np.float64).reshape(-1)
        if not np.isscalar(wv) and wv.shape[0] != x.shape[0]:
            raise ValueError("DimensionMismatch: x and w differ in length")
        self.handle.field_update(x, wv, self.chi)

    def efield(self, E, x):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        out = self.handle.field_eval(x)
        E[...] = out.reshape(np.shape(E))
        return E

    def energy(self):
        return self.handle.field_energy()

    def coefficients(self):
        return self.handle.field_coefficients()

    def __call__(self, E, x, w, t=0.0):
        self.update(x, w, t)
        return self.efield(E, x)


B200PoissonField = ScaledPoissonField
