"""Reduced online model with DEIM -- host API kept from the reference, compute in librbm_b200.

  ReducedElectricField / DEIMElectricField ...... src/particles/electric_field.jl:2-43
  ReducedIntegratorCache ........................ src/particles/time_marching.jl:66-79
  reduced_integrate_vp .......................... src/particles/time_marching.jl:108-150
  driver ........................................ scripts/bump_on_tail_3_testing.jl:108-138

The reference integrates one parameter sample per call; here all samples advance in lock-step on the GPU
(``reduced_integrate_vp_all``), and the per-sample call is the same thing with one sample.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .integrator import _handle_for, _load_particles
from .particles import ParticleList
from .poisson import PoissonSolverPBSplines
from .reducedbasis import _colmajor
from .snapshots import IntegratorParameters, Snapshots


class DEIMElectricField:
    """``DEIMElectricField(field, Psi_x, Psi_e, Pi_e, X)``: the three reduced operators Pi'Psi_x and
    Psi_x'Psi_e (Pi'Psi_e)^-1 are formed on the device when the model is bound to the particles."""

    def __init__(self, poisson: PoissonSolverPBSplines, Psi_x, Psi_e, Pi_e):
        self.poisson = poisson
        self.Psi_x, self.N, self.kp, self.ldx = _colmajor(Psi_x)
        self.Psi_e, n2, self.ke, self.lde = _colmajor(Psi_e)
        if n2 != self.N:
            raise ValueError("DimensionMismatch: Psi_x and Psi_e differ in row count")
        Pi_e = np.asarray(Pi_e)
        if Pi_e.ndim == 2:                       # dense N x k_e 0/1 matrix as the reference stores it
            if Pi_e.shape != (self.N, self.ke):
                raise ValueError("DimensionMismatch: Pi_e must be N x k_e")
            idx = np.argmax(Pi_e, axis=0)
        else:
            idx = Pi_e
        self.idx = np.ascontiguousarray(idx, dtype=np.int64)
        if self.idx.shape != (self.ke,):
            raise ValueError("DimensionMismatch: one DEIM row per column of Psi_e")


class ReducedModel:
    """rbm_reduced: a DEIMElectricField bound to the particles (x0, v0, w) on one GPU."""

    def __init__(self, particles: ParticleList, rfield: DEIMElectricField, ctx=None):
        if len(particles) != rfield.N:
            raise ValueError("DimensionMismatch: length(particles) != size(Psi, 1)")
        self.rfield = rfield
        self.handle = _handle_for(len(particles), rfield.poisson, ctx)
        self.ctx = self.handle.ctx
        _load_particles(self.handle, particles)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.rbm_reduced_create(self.handle.h, _lib.ptr(rfield.Psi_x), rfield.ldx, rfield.kp,
                                                       _lib.ptr(rfield.Psi_e), rfield.lde, rfield.ke,
                                                       rfield.idx.ctypes.data, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.rbm_reduced_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, chis, dt, nt, ns, Zx=None, Zv=None, X=None, V=None, Phi=None, W=None, K=None, M=None):
        chis = np.ascontiguousarray(chis, dtype=np.float64).reshape(-1)
        nparam, rf = chis.shape[0], self.rfield
        for name, a in (("Zx", Zx), ("Zv", Zv)):
            _lib.check_buffer(a, nparam * int(ns) * rf.kp, name)
        for name, a in (("X", X), ("V", V)):
            _lib.check_buffer(a, nparam * int(ns) * rf.N, name)
        _lib.check_buffer(Phi, nparam * int(ns) * rf.poisson.nh, "Phi")
        for name, a in (("W", W), ("K", K), ("M", M)):
            _lib.check_buffer(a, nparam * int(ns), name)
        self.ctx.check(self.ctx.lib.rbm_reduced_run(self.h, chis.ctypes.data, chis.shape[0], float(dt), int(nt), int(ns),
                                                    *[_lib.ptr(a) for a in (Zx, Zv, X, V, Phi, W, K, M)]))


def reduced_integrate_vp_all(particles: ParticleList, rfield: DEIMElectricField, chis, IP: IntegratorParameters,
                             RSS: Snapshots | None = None, ctx=None, return_z: bool = False):
    """All samples in lock-step; fills Snapshots X, V, Phi, W, K, M (A stays zero, as in the reference's save_solution)."""
    chis = np.ascontiguousarray(chis, dtype=np.float64).reshape(-1)
    if IP.nparam != chis.shape[0] or IP.np != len(particles):
        raise ValueError("DimensionMismatch between IntegratorParameters, particles and chi")
    RSS = RSS or Snapshots.from_integrator(IP)
    model = ReducedModel(particles, rfield, ctx)
    Zx = np.zeros((IP.nparam, IP.ns, rfield.kp)) if return_z else None
    Zv = np.zeros((IP.nparam, IP.ns, rfield.kp)) if return_z else None
    model.run(chis, IP.dt, IP.nt, IP.ns, Zx=Zx, Zv=Zv, X=RSS.X, V=RSS.V, Phi=RSS.Phi, W=RSS.W, K=RSS.K, M=RSS.M)
    model.close()
    return (RSS, Zx, Zv) if return_z else RSS


def reduced_integrate_vp(particles: ParticleList, Psi, params: dict, rfield: DEIMElectricField, RSS: Snapshots,
                         IP: IntegratorParameters, p: int, ctx=None):
    """reduced_integrate_vp(P0, Psi, params, efield, SS, IP, IC, p): one sample, written into slot p of SS."""
    one = IntegratorParameters(IP.dt, IP.nt, IP.ns, IP.nh, IP.np, 1)
    ss = reduced_integrate_vp_all(particles, rfield, [params["chi"]], one, ctx=ctx)
    for k in ("X", "V", "Phi", "W", "K", "M"):
        getattr(RSS, k)[p] = getattr(ss, k)[0]
    return RSS
