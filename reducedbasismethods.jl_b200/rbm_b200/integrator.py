"""VPIntegratorParameters / VPIntegratorCache / integrate_vp! -- the full-order integrator of
VlasovMethods.jl as the reference drives it (scripts/bump_on_tail_1_training.jl:67,70,97,100-108),
backed by librbm_b200.  Two entry points:

* ``integrate_vp(particles, efield, lparams, IP, IC, save=True)`` -- the upstream per-sample
  signature, a drop-in for the unmodified script loop (fills IC.X, V, A, Phi, W, K, M);
* ``integrate_vp_all(particles, poisson, chis, IP, SS)`` -- all parameter samples advanced in
  lock-step on the GPU, written straight into ``Snapshots`` (the sample loop of
  scripts/bump_on_tail_1_training.jl:87-109 moved inside the library).
"""
from __future__ import annotations

import numpy as np

from .particles import ParticleList
from .poisson import PICHandle, PoissonSolverPBSplines, ScaledPoissonField
from .snapshots import IntegratorParameters, Snapshots


class VPIntegratorParameters:
    """VPIntegratorParameters(dt, nt, ns, nh, np)  (scripts/bump_on_tail_1_training.jl:67)."""

    def __init__(self, dt, nt, ns, nh, np_):
        self.dt = float(dt); self.nt = int(nt); self.ns = int(ns); self.nh = int(nh); self.np = int(np_)
        self.t = np.linspace(0.0, self.dt * self.nt, self.ns)

    def with_paramspace(self, pspace) -> IntegratorParameters:
        """IntegratorParameters(ip, pspace)  (scripts/bump_on_tail_1_training.jl:14-16)."""
        return IntegratorParameters(self.dt, self.nt, self.ns, self.nh, self.np, len(pspace))


class VPIntegratorCache:
    """X, V, A (np x ns), Phi (nh x ns), W, K, M (ns)  (scripts/bump_on_tail_1_training.jl:70,100-108);
    stored C-order (ns, np) == Julia column-major (np, ns)."""

    def __init__(self, IP: VPIntegratorParameters):
        self.X = np.zeros((IP.ns, IP.np)); self.V = np.zeros((IP.ns, IP.np)); self.A = np.zeros((IP.ns, IP.np))
        self.Phi = np.zeros((IP.ns, IP.nh))
        self.W = np.zeros(IP.ns); self.K = np.zeros(IP.ns); self.M = np.zeros(IP.ns)


_handles: dict = {}


def _handle_for(np_, poisson: PoissonSolverPBSplines, ctx=None) -> PICHandle:
    key = (np_, poisson.p, poisson.nh, poisson.L, id(ctx))
    h = _handles.get(key)
    if h is None or h.h is None or getattr(h.ctx, "h", None) is None:
        h = PICHandle(np_, poisson, ctx)
        _handles[key] = h
    return h


def _load_particles(handle: PICHandle, particles: ParticleList):
    if particles.x.shape[0] != 1:
        raise ValueError("only nd = 1 (1D1V) is implemented, as in every reference script")
    wu = particles.uniform_weight()
    handle.set_particles(particles.vec_x(), particles.vec_v(), wu if wu is not None else particles.vec_w())


def integrate_vp(particles: ParticleList, efield: ScaledPoissonField, lparams: dict, IP: VPIntegratorParameters,
                 IC: VPIntegratorCache | None = None, save: bool = True, ctx=None):
    """integrate_vp!(particles, efield, lparams, IP, IC; save) for ONE sample; chi is taken from the
    field like upstream (ScaledPoissonField carries it) and must agree with lparams['chi'] if given.
    `particles` is left untouched (the reference re-uses it for every sample)."""
    if len(particles) != IP.np:
        raise ValueError("DimensionMismatch: length(particles) != IP.np")
    chi = efield.chi
    if lparams and "chi" in lparams and float(lparams["chi"]) != chi:
        raise ValueError("lparams.chi differs from the chi of the field")
    IC = IC or VPIntegratorCache(IP)
    h = _handle_for(IP.np, efield.poisson, ctx)
    _load_particles(h, particles)
    if save:
        h.run([chi], IP.dt, IP.nt, IP.ns, IC.X, IC.V, IC.A, IC.Phi, IC.W, IC.K, IC.M)
    else:
        h.run([chi], IP.dt, IP.nt, 2)
    return IC


def integrate_vp_all(particles: ParticleList, poisson: PoissonSolverPBSplines, chis, IP: IntegratorParameters,
                     SS: Snapshots | None = None, ctx=None, device=None) -> Snapshots:
    """All samples at once, straight into Snapshots (same memory layout as the device buffers).  With device
    snapshots (SS.on_device, or device="cuda" when SS is None) nothing leaves HBM, and the call only QUEUES the work on the
    context's stream (a blocking stream: torch's default stream is ordered after it; call ctx.synchronize() before handing
    the tensors to code on another non-blocking stream)."""
    chis = np.ascontiguousarray(chis, dtype=np.float64).reshape(-1)
    if IP.nparam != chis.shape[0] or IP.np != len(particles) or IP.nh != poisson.nh:
        raise ValueError("DimensionMismatch between IntegratorParameters, particles, poisson and chi")
    SS = SS or Snapshots.from_integrator(IP, device=device)
    h = _handle_for(IP.np, poisson, ctx)
    _load_particles(h, particles)
    h.run(chis, IP.dt, IP.nt, IP.ns, SS.X, SS.V, SS.A, SS.Phi, SS.W, SS.K, SS.M)
    return SS
