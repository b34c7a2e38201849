"""Initial-condition samplers: the host draw (input producer of the parity tests) and the device draw.

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


def draw_accept_reject_device(n: int, params: dict | None = None, seed: int = 1234, ctx=None, first: int = 0,
                              total: int | None = None, device_out: bool = True):
    """The same distribution drawn ON THE DEVICE (`rbm_sample_bump_on_tail`, counter-based Philox4x32-10): particle
    p = first + i depends only on (seed, p), so ranks holding particle chunks draw disjoint pieces of ONE global
    sample (`total` = global particle count, default n).  Returns (x, v, w) as torch CUDA tensors (device_out) or
    NumPy arrays, plus nothing else; the number of rejected candidates is available as
    ``draw_accept_reject_device.last_rejected``."""
    import ctypes as C

    from . import _lib
    from .engine import default_context

    ctx = ctx or default_context()
    p = dict(REF_PARAMS)
    if params:
        p.update(params)
    L = domain_length(p["kappa"])
    total = n if total is None else int(total)
    if device_out:
        import torch
        dev = f"cuda:{ctx.device}" if hasattr(ctx, "device") else "cuda"
        x = torch.empty(n, dtype=torch.float64, device=dev)
        v = torch.empty(n, dtype=torch.float64, device=dev)
        w = torch.empty(n, dtype=torch.float64, device=dev)
    else:
        x = np.empty(n); v = np.empty(n); w = np.empty(n)
    rej = C.c_int64(0)
    ctx.check(ctx.lib.rbm_sample_bump_on_tail(ctx.h, int(n), int(first), int(seed) & (2 ** 64 - 1), float(p["kappa"]),
                                              float(p["eps"]), float(p["a"]), float(p["v0"]), float(p["sigma"]),
                                              L / total, _lib.ptr(x), _lib.ptr(v), _lib.ptr(w), C.byref(rej)))
    draw_accept_reject_device.last_rejected = rej.value
    return x, v, w
