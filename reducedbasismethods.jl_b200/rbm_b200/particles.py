"""ParticleList -- host container kept from ParticleMethods.jl as the reference uses it:
``ParticleList(x, v, w)`` with x, v of shape (nd, np) and w of shape (1, np)
(test/trainingset_tests.jl:17-21), fields .x .v .w .list (src/particles/time_marching.jl:122-124,
test/trainingset_tests.jl:63) and ``length`` = np (src/trainingset.jl:21).  On the device the
particles are structure-of-arrays; the host container only has to hand over contiguous
vectors (``vec(P.x)``)."""
from __future__ import annotations

import numpy as np


class ParticleList:
    def __init__(self, x, v, w):
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        v = np.atleast_2d(np.asarray(v, dtype=np.float64))
        w = np.atleast_2d(np.asarray(w, dtype=np.float64))
        if x.shape != v.shape or w.shape != (1, x.shape[1]):
            raise ValueError("DimensionMismatch: x, v must be (nd, np) and w (1, np)")
        # one stacked array with x, v, w as views, like the upstream type
        self.list = np.concatenate([x, v, w], axis=0)
        nd = x.shape[0]
        self.x = self.list[:nd]
        self.v = self.list[nd:2 * nd]
        self.w = self.list[2 * nd:]

    def __len__(self):
        return self.list.shape[1]

    def __eq__(self, o):
        return isinstance(o, ParticleList) and np.array_equal(self.list, o.list)

    def vec_x(self):
        return np.ascontiguousarray(self.x.reshape(-1))

    def vec_v(self):
        return np.ascontiguousarray(self.v.reshape(-1))

    def vec_w(self):
        return np.ascontiguousarray(self.w.reshape(-1))

    def uniform_weight(self):
        """The common weight if all weights are bit-identical (the sampler's w = L/np), else None."""
        w = self.w.reshape(-1)
        return float(w[0]) if np.all(w == w[0]) else None
