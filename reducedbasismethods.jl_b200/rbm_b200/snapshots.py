"""Snapshots / IntegratorParameters -- containers kept from the reference
(src/particles/snapshots.jl:2-79, src/particles/time_marching.jl:7-64).

Julia arrays are column-major ``(nd, np, ns, nparam)``; the NumPy mirror stores the same
memory as C-order ``(nparam, ns, np, nd)`` so that the buffer handed to the C ABI is
byte-identical to the Julia one (particle index fastest)."""
from __future__ import annotations

import numpy as np


class IntegratorParameters:
    """dt, nt, ns, nh, np, nparam and t = range(0, dt*nt, length=ns) (time_marching.jl:7-22)."""

    def __init__(self, dt, nt, ns, nh, np_, nparam):
        self.dt = float(dt)
        self.nt = int(nt)
        self.ns = int(ns)
        self.nh = int(nh)
        self.np = int(np_)
        self.nparam = int(nparam)
        self.t = np.linspace(0.0, self.dt * self.nt, self.ns)

    def __eq__(self, o):
        return (isinstance(o, IntegratorParameters)
                and (self.dt, self.nt, self.ns, self.nh, self.np, self.nparam)
                == (o.dt, o.nt, o.ns, o.nh, o.np, o.nparam) and np.array_equal(self.t, o.t))

    def attrs(self):
        """HDF5 attribute names of the reference (time_marching.jl:56-64)."""
        return dict(dt=self.dt, nt=self.nt, ns=self.ns, nh=self.nh, np=self.np, nparam=self.nparam)

    @classmethod
    def from_attrs(cls, d):
        return cls(d["dt"], d["nt"], d["ns"], d["nh"], d["np"], d["nparam"])


class Snapshots:
    """X, V, A: (nparam, ns, np, nd); Phi: (nparam, ns, nh, nd); W, K, M: (nparam, ns)
    == Julia (nd, np, ns, nparam) / (nd, nh, ns, nparam) / (ns, nparam) (snapshots.jl:16-25)."""

    FIELDS = ("X", "V", "A", "Phi", "W", "K", "M")

    def __init__(self, nd=None, np_=None, nh=None, nt=None, nparam=None, dtype=np.float64, arrays=None):
        if arrays is not None:
            for k in self.FIELDS:
                setattr(self, k, arrays[k])
            return
        self.X = np.zeros((nparam, nt, np_, nd), dtype=dtype)
        self.V = np.zeros((nparam, nt, np_, nd), dtype=dtype)
        self.A = np.zeros((nparam, nt, np_, nd), dtype=dtype)
        self.Phi = np.zeros((nparam, nt, nh, nd), dtype=dtype)
        self.W = np.zeros((nparam, nt), dtype=dtype)
        self.K = np.zeros((nparam, nt), dtype=dtype)
        self.M = np.zeros((nparam, nt), dtype=dtype)

    @classmethod
    def from_integrator(cls, ip: IntegratorParameters, dtype=np.float64):
        return cls(1, ip.np, ip.nh, ip.ns, ip.nparam, dtype)      # snapshots.jl:27-29

    def __eq__(self, o):
        return isinstance(o, Snapshots) and all(np.array_equal(getattr(self, k), getattr(o, k)) for k in self.FIELDS)

    def matrix(self, name):
        """reshape(SS.X, (np, ns*nparam)) of src/algorithms/evd.jl:124-126, as a Fortran-order view."""
        a = getattr(self, name)
        nparam, ns, n, nd = a.shape
        return a.reshape(nparam * ns, n * nd).T          # (np, ns*nparam), column-major memory

    def to_dict(self):
        return {k: getattr(self, k) for k in self.FIELDS}
