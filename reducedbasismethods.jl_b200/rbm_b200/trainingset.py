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
        d = {}
        for k, v in self.parameters.items():
            d[f"parameters/{k}"] = np.asarray(v)
        ps = self.paramspace.to_dict()
        for k, v in ps["samples"].items():
            d[f"parameterspace/samples/{k}"] = v
        for k, v in ps["parameters"].items():
            d[f"parameterspace/parameters/{k}/minimum"] = np.asarray(v["minimum"])
            d[f"parameterspace/parameters/{k}/maximum"] = np.asarray(v["maximum"])
            d[f"parameterspace/parameters/{k}/samples"] = np.asarray(v["samples"])
        d["initial_conditions/list"] = self.initconds.list
        for k, v in self.snapshots.to_dict().items():
            d[f"snapshots/{k}"] = v
        for k, v in self.integrator.attrs().items():
            d[f"integrator/{k}"] = np.asarray(v)
        d["p"] = np.asarray(self.poisson.p)
        d["L"] = np.asarray(self.poisson.L)
        np.savez(path, **d)

    @classmethod
    def load(cls, path):
        z = np.load(path)
        params = {k.split("/", 1)[1]: float(z[k]) for k in z.files if k.startswith("parameters/")}
        names = []
        for k in z.files:
            if k.startswith("parameterspace/samples/"):
                names.append(k.rsplit("/", 1)[1])
        psd = {"samples": {n: z[f"parameterspace/samples/{n}"] for n in names},
               "parameters": {n: {"minimum": float(z[f"parameterspace/parameters/{n}/minimum"]),
                                  "maximum": float(z[f"parameterspace/parameters/{n}/maximum"]),
                                  "samples": z[f"parameterspace/parameters/{n}/samples"]} for n in names}}
        pl = z["initial_conditions/list"]
        nd = (pl.shape[0] - 1) // 2
        particles = ParticleList(pl[:nd], pl[nd:2 * nd], pl[2 * nd:])
        ss = Snapshots(arrays={k: z[f"snapshots/{k}"] for k in Snapshots.FIELDS})
        ip = IntegratorParameters.from_attrs({k: z[f"integrator/{k}"].item() for k in ("dt", "nt", "ns", "nh", "np", "nparam")})
        poisson = PoissonSolverPBSplines(int(z["p"]), ip.nh, float(z["L"]))   # nh from integrator attrs (poisson.jl:11-21)
        return cls(params, ParameterSpace.from_dict(psd), particles, ss, ip, poisson)
