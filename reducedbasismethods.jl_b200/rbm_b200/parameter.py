"""Parameter / CartesianParameterSampler / ParameterSpace -- host API kept from the reference
(src/parameter.jl, src/parametersampler.jl, src/parameterspace.jl); pure host code."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np


class Parameter:
    """src/parameter.jl:14-34: named scalar range plus sorted unique samples (or None)."""

    def __init__(self, name, minimum, maximum=None, samples=None):
        # Parameter(name, samples::AbstractVector)  (parameter.jl:34)
        if maximum is None and samples is None and not np.isscalar(minimum):
            samples = np.asarray(minimum, dtype=np.float64)
            minimum, maximum = float(samples.min()), float(samples.max())
        # Parameter(name, min, max, n::Int) -> LinRange(min, max, n)  (parameter.jl:33)
        if isinstance(samples, (int, np.integer)) and not isinstance(samples, bool):
            samples = np.linspace(minimum, maximum, int(samples))
        assert minimum <= maximum                                        # parameter.jl:21
        if samples is not None:
            samples = np.asarray(samples, dtype=np.float64).reshape(-1)
            assert np.all(samples >= minimum) and np.all(samples <= maximum)   # parameter.jl:23-24
            samples = np.unique(samples)                                  # _sort: sort(unique(...)) (:8)
        self.name = str(name)
        self.minimum = float(minimum)
        self.maximum = float(maximum)
        self.samples = samples

    def __eq__(self, o):
        return (isinstance(o, Parameter) and self.name == o.name and self.minimum == o.minimum
                and self.maximum == o.maximum
                and ((self.samples is None and o.samples is None)
                     or (self.samples is not None and o.samples is not None
                         and np.array_equal(self.samples, o.samples))))

    def __hash__(self):
        return hash((self.name, self.minimum, self.maximum,
                     None if self.samples is None else self.samples.tobytes()))

    def __len__(self):                       # parameter.jl:59-60
        return 0 if self.samples is None else len(self.samples)

    @property
    def size(self):
        return (len(self),)

    def collect(self):
        return self.samples

    def hassamples(self):                    # parameter.jl:65-66
        return len(self) > 0

    def __getitem__(self, i):                # parameter.jl:68-71
        if self.samples is None:
            raise IndexError(f"Parameter {self.name} indexed with {i} but has no samples.")
        return self.samples[i]


def named_parameters(*parameters):
    """Base.NamedTuple(parameters...)  (parameter.jl:73-76)."""
    return OrderedDict((p.name, p) for p in parameters)


class CartesianParameterSampler:
    pass


def sample(sampler, *parameters):
    """src/parametersampler.jl:20-36: tensor-product samples, FIRST parameter fastest
    (CartesianIndices order; pinned by test/parameterspace_tests.jl:2-7)."""
    if len(parameters) == 1 and isinstance(parameters[0], dict):
        parameters = tuple(parameters[0].values())
    for p in parameters:
        assert p.hassamples()
    shape = [len(p) for p in parameters]
    total = int(np.prod(shape))
    table = OrderedDict()
    stride = 1
    for p, n in zip(parameters, shape):
        idx = (np.arange(total) // stride) % n
        table[p.name] = p.samples[idx]
        stride *= n
    return table


class ParameterSpace:
    """src/parameterspace.jl:6-36: parameters plus a table of samples; ps(i) -> dict of values."""

    def __init__(self, *args):
        if len(args) == 2 and isinstance(args[0], dict) and isinstance(args[1], dict):
            self.parameters, self.samples = OrderedDict(args[0]), OrderedDict(args[1])
            return
        if args and isinstance(args[0], CartesianParameterSampler):
            sampler, params = args[0], args[1:]
        else:
            sampler, params = CartesianParameterSampler(), args
        if len(params) == 1 and isinstance(params[0], dict):
            params = tuple(params[0].values())
        self.parameters = named_parameters(*params)
        self.samples = sample(sampler, *params)

    def __call__(self, i):                   # parameterspace.jl:36 (0-based here)
        return {k: float(self.samples[k][i]) for k in self.parameters}

    def __len__(self):
        return len(next(iter(self.samples.values()))) if self.samples else 0

    def ndims(self):
        return len(self.parameters)

    @property
    def size(self):
        return (len(self), len(self.parameters))

    def eachindex(self):
        return range(len(self))

    def __getitem__(self, i):
        return tuple(self.samples[k][i] for k in self.samples)

    def column(self, name):
        return self.samples[name]

    def __eq__(self, o):
        return (isinstance(o, ParameterSpace) and list(self.parameters.items()) == list(o.parameters.items())
                and list(self.samples) == list(o.samples)
                and all(np.array_equal(self.samples[k], o.samples[k]) for k in self.samples))

    # persistence (the reference uses HDF5 groups "samples" / "parameters": parameterspace.jl:71-83)
    def to_dict(self):
        return {"samples": {k: np.asarray(v) for k, v in self.samples.items()},
                "parameters": {k: {"minimum": p.minimum, "maximum": p.maximum, "samples": p.samples}
                               for k, p in self.parameters.items()}}

    @classmethod
    def from_dict(cls, d):
        params = OrderedDict((k, Parameter(k, v["minimum"], v["maximum"], v["samples"]))
                             for k, v in d["parameters"].items())
        return cls(params, OrderedDict((k, np.asarray(v)) for k, v in d["samples"].items()))
