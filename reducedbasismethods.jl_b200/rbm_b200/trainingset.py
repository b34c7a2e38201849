"""TrainingSet -- bundle kept from the reference (src/trainingset.jl:2-69)."""
from __future__ import annotations

import numpy as np

from .parameter import ParameterSpace
from .particles import ParticleList
from .poisson import PoissonSolverPBSplines
from .snapshots import IntegratorParameters, Snapshots


class TrainingSet:
    def __init__(self, parameters: dict, paramspace: ParameterSpace, initconds: ParticleList, snapshots: Snapshots,
                 integrator: IntegratorParameters, poisson: PoissonSolverPBSplines):
        self.parameters = dict(parameters)
        self.paramspace = paramspace
        self.initconds = initconds
        self.snapshots = snapshots
        self.integrator = integrator
        self.poisson = poisson

    @classmethod
    def create(cls, particles: ParticleList, poisson: PoissonSolverPBSplines, nt: int, parameters: dict,
               pspace: ParameterSpace, ip: IntegratorParameters, nd: int = 1, device=None):
        """TrainingSet(particles, poisson, nt, parameters, pspace, ip)  (trainingset.jl:24-26 -> :20-22 -> :15-18).
        device="cuda": the snapshots live in HBM and are handed to the POD stage without a host round trip."""
        ss = Snapshots(nd, len(particles), len(poisson), nt, len(pspace), device=device)
        return cls(parameters, pspace, particles, ss, ip, poisson)

    def __eq__(self, o):
        return (isinstance(o, TrainingSet) and self.parameters == o.parameters and self.paramspace == o.paramspace
                and self.initconds == o.initconds and self.snapshots == o.snapshots
                and self.integrator == o.integrator and self.poisson == o.poisson)

    # persistence: the reference writes HDF5 groups parameters / parameterspace / initial_conditions /
    # snapshots / integrator and attrs p, L (trainingset.jl:56-65).  There is no HDF5 toolchain in this
    # image, so the same tree is written as a NumPy .npz with '/'-separated keys.
    def save(self, path):
        from .reducedbasis import _tree_of

        d = _tree_of(self.parameters, self.paramspace, self.initconds, self.integrator, self.poisson)
        for k, v in self.snapshots.to_dict().items():
            d[f"snapshots/{k}"] = v
        np.savez(path, **d)

    @classmethod
    def load(cls, path):
        from .reducedbasis import _tree_from

        z = np.load(path)
        params, pspace, particles, ip, poisson = _tree_from(z)
        ss = Snapshots(arrays={k: z[f"snapshots/{k}"] for k in Snapshots.FIELDS})
        return cls(params, pspace, particles, ss, ip, poisson)
