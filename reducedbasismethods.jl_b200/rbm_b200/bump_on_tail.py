"""Host-side initial-condition sampler (input producer; stays on the host).

Mirrors ``VlasovMethods.BumpOnTail.draw_accept_reject(np, params)`` as it is called at
reference scripts/bump_on_tail_1_training.jl:79 (after ``Random.seed!(1234)``, :76):
accept-reject draw from the bump-on-tail distribution

    f(x, v) = (1 + eps*cos(kappa*x)) * [ (1-a) N(0,1)(v) + a N(v0, sigma^2)(v) ],   x in [0, L), L = 2*pi/kappa

with uniform weights ``w = L/np`` (notebooks/VP-ROM-Projections.ipynb cell 2).  Julia's
MersenneTwister stream cannot be reproduced here, so the draw is seeded with NumPy's
PCG64 (``default_rng(seed)``); inputs cross the C-ABI as arrays, so the stream is not
part of the parity contract.
"""
from __future__ import annotations

import math

import numpy as np

REF_PARAMS = dict(kappa=0.3, eps=0.03, a=0.1, v0=4.5, sigma=0.5, chi=1.0)


def domain_length(kappa: float) -> float:
    """L = 2*pi/kappa  (scripts/bump_on_tail_1_training.jl:64)."""
    return 2.0 * math.pi / kappa


def draw_accept_reject(n: int, params: dict | None = None, seed: int = 1234):
    """Returns (x, v, w) float64 arrays of length n."""
    p = dict(REF_PARAMS)
    if params:
        p.update(params)
    L = domain_length(p["kappa"])
    rng = np.random.default_rng(seed)
    # v: exact draw from the two-Maxwellian mixture
    tail = rng.random(n) < p["a"]
    v = rng.standard_normal(n)
    v = np.where(tail, p["v0"] + p["sigma"] * v, v)
    # x: accept-reject against the envelope (1 + eps)
    x = np.empty(n, dtype=np.float64)
    filled = 0
    while filled < n:
        m = max(1024, int(1.1 * (n - filled)))
        cand = rng.random(m) * L
        u = rng.random(m) * (1.0 + p["eps"])
        ok = cand[u <= 1.0 + p["eps"] * np.cos(p["kappa"] * cand)]
        take = min(ok.shape[0], n - filled)
        x[filled:filled + take] = ok[:take]
        filled += take
    w = np.full(n, L / n, dtype=np.float64)
    return x, v, w
