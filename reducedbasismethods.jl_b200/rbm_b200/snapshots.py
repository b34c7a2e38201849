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

    def __init__(self, nd=None, np_=None, nh=None, nt=None, nparam=None, dtype=np.float64, arrays=None, device=None):
        """device=None: NumPy host arrays (what the reference's HDF5 round trip holds).  device="cuda[:i]": the same
        buffers as torch tensors in HBM -- stage 1 (integrate_vp_all) writes them and stage 2 (get_PODBasis_*, project,
        ReducedBasis.from_trainingset) reads them through the C ABI's device-pointer path without a host copy
        (the reference passes them through HDF5: src/trainingset.jl:37-65)."""
        if arrays is not None:
            for k in self.FIELDS:
                setattr(self, k, arrays[k])
            return
        if device is None:
            z = lambda *s: np.zeros(s, dtype=dtype)
        else:
            import torch
            if np.dtype(dtype) != np.float64:
                raise TypeError("device snapshots are float64 (the compute type of the path)")
            z = lambda *s: torch.zeros(s, dtype=torch.float64, device=device)
        self.X = z(nparam, nt, np_, nd)
        self.V = z(nparam, nt, np_, nd)
        self.A = z(nparam, nt, np_, nd)
        self.Phi = z(nparam, nt, nh, nd)
        self.W = z(nparam, nt)
        self.K = z(nparam, nt)
        self.M = z(nparam, nt)

    @classmethod
    def from_integrator(cls, ip: IntegratorParameters, dtype=np.float64, device=None):
        return cls(1, ip.np, ip.nh, ip.ns, ip.nparam, dtype, device=device)      # snapshots.jl:27-29

    @property
    def on_device(self) -> bool:
        return not isinstance(self.X, np.ndarray)

    def to_host(self, pinned: bool = False, non_blocking: bool = False) -> "Snapshots":
        """Host copy (only needed for persistence).  pinned + non_blocking: the copies are queued on the current torch
        stream and overlap whatever stage 2 launches next; synchronise before touching the result."""
        if not self.on_device:
            return self
        import torch
        out = {}
        for k in self.FIELDS:
            a = getattr(self, k)
            h = torch.empty(a.shape, dtype=a.dtype, device="cpu", pin_memory=pinned)
            h.copy_(a, non_blocking=non_blocking and pinned)
            out[k] = h.numpy()
        return Snapshots(arrays=out)

    def to_device(self, device="cuda") -> "Snapshots":
        if self.on_device:
            return self
        import torch
        return Snapshots(arrays={k: torch.from_numpy(np.ascontiguousarray(getattr(self, k))).to(device) for k in self.FIELDS})

    def __eq__(self, o):
        if not isinstance(o, Snapshots):
            return False
        a, b = self.to_host(), o.to_host()
        return all(np.array_equal(getattr(a, k), getattr(b, k)) for k in self.FIELDS)

    def matrix(self, name):
        """reshape(SS.X, (np, ns*nparam)) of src/algorithms/evd.jl:124-126, as a column-major view (host or device)."""
        a = getattr(self, name)
        nparam, ns, n, nd = a.shape
        return a.reshape(nparam * ns, n * nd).T          # (np, ns*nparam), column-major memory

    def to_dict(self):
        h = self.to_host()
        return {k: getattr(h, k) for k in self.FIELDS}
