"""oracle/reduced_numpy.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy restatement of the reference's reduced online model with DEIM -- all of it is in the tree:
  ReducedElectricField / DEIMElectricField .... src/particles/electric_field.jl:2-43
  ReducedIntegratorCache, save_solution ....... src/particles/time_marching.jl:66-105
  reduced_integrate_vp ........................ src/particles/time_marching.jl:108-150
  driver ...................................... scripts/bump_on_tail_3_testing.jl:108-138
The wrapped full field (update!/efield!/energy/coefficients of ScaledPoissonField) is oracle/pic_numpy.py
(that part is external to the tree; see its header).  No reference test covers this path; it is pinned by being a
literal restatement of in-tree code.

Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

from . import pic_numpy as pn


class ScaledField:
    """ScaledPoissonField(poisson, chi): update!/efield!/energy/coefficients."""

    def __init__(self, L, nh, chi):
        self.L, self.nh, self.chi = L, nh, chi
        self.phi = np.zeros(nh)
        self.W = 0.0

    def update(self, x, w):
        self.phi, self.W = pn.poisson(pn.deposit(x, w, self.L, self.nh), self.nh, self.L, self.chi)

    def efield(self, x):
        return pn.gather(self.phi, x, self.L, self.nh)


class ReducedField:
    """ReducedElectricField(field, Psi_x, Psi_p, Psi_e, X)  (electric_field.jl:2-17): three matrices around a field."""

    def __init__(self, field, Psi_x, Psi_p, Psi_e):
        self.field, self.Psi_x, self.Psi_p, self.Psi_e = field, Psi_x, Psi_p, Psi_e

    def update(self, Z, w):                                  # electric_field.jl:22-25
        self.field.update(self.Psi_x @ Z, w)

    def efield(self, Z):                                     # electric_field.jl:27-31
        y = self.Psi_p @ Z
        e = self.field.efield(y)
        return self.Psi_e @ e


def deim_field(field, Psi_x, Psi_e, idx):
    """DEIMElectricField(field, Psi_x, Psi_e, Pi_e, X)  (electric_field.jl:39-43) with Pi_e given by its rows idx."""
    PiT_Psix = Psi_x[idx, :]                                 # Pi_e' * Psi_x
    PsixT_Pe = Psi_x.T @ Psi_e @ np.linalg.inv(Psi_e[idx, :])
    return ReducedField(field, Psi_x, PiT_Psix, PsixT_Pe)


def reduced_integrate(x0, v0, w, Psi, rfield: ReducedField, chi, dt, nt, ns):
    """reduced_integrate_vp (time_marching.jl:108-150) for one sample.  Returns Zx, Zv (ns, k), X, V (ns, N),
    Phi (ns, nh), W, K, M (ns)."""
    N, k = Psi.shape
    nsave = nt // (ns - 1)
    w = np.broadcast_to(np.asarray(w, dtype=np.float64), (N,))
    zx = Psi.T @ x0                                          # :122
    zv = Psi.T @ v0                                          # :123
    dte = dt * chi                                           # :127
    out = dict(Zx=np.zeros((ns, k)), Zv=np.zeros((ns, k)), X=np.zeros((ns, N)), V=np.zeros((ns, N)),
               Phi=np.zeros((ns, rfield.field.nh)), W=np.zeros(ns), K=np.zeros(ns), M=np.zeros(ns))
    ts = 0

    def save(step):
        nonlocal ts
        if step % nsave == 0 and ts < ns:                    # save_solution :81-105
            x = Psi @ zx
            v = Psi @ zv
            rfield.update(zx, w)
            out["Zx"][ts] = zx; out["Zv"][ts] = zv; out["X"][ts] = x; out["V"][ts] = v
            out["Phi"][ts] = rfield.field.phi
            out["W"][ts] = rfield.field.W
            out["K"][ts] = float(np.sum((v * w * v).astype(np.longdouble)) / 2)
            out["M"][ts] = float(np.sum((w * v).astype(np.longdouble)))
            ts += 1

    save(0)                                                  # :130
    for it in range(1, nt + 1):
        zx = zx + 0.5 * dte * zv                             # :137
        rfield.update(zx, w)                                 # :140  efield(z_a, z_x, w, t) = update! + efield!
        za = rfield.efield(zx)
        zv = zv + dte * za                                   # :143
        zx = zx + 0.5 * dte * zv                             # :146
        save(it)
    return out
