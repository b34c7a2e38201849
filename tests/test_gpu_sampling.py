"""GPU tests of the device sampler and growth-rate regression through the C ABI against their
NumPy restatements (oracle/sampler_numpy.py, oracle/regression_numpy.py)."""
import numpy as np
import pytest

from oracle import regression_numpy as reg
from oracle import sampler_numpy as smp

pytestmark = pytest.mark.gpu

import rbm_b200 as r  # noqa: E402


def test_sampler_matches_oracle(ctx):
    n = 100_000
    x, v, w = r.BumpOnTail.draw_accept_reject_device(n, seed=1234, ctx=ctx, device_out=False)
    xo, vo, rej = smp.sample_bump_on_tail(n, seed=1234)
    L = r.BumpOnTail.domain_length(0.3)
    assert np.array_equal(x, xo)                                  # uniforms are exact; accept decisions identical
    assert np.max(np.abs(v - vo)) < 1e-13                         # log/cos/sqrt of two math libraries
    assert np.all(w == L / n)
    assert r.BumpOnTail.draw_accept_reject_device.last_rejected == rej
    # another seed and other parameters
    prm = dict(kappa=0.5, eps=0.2, a=0.3, v0=3.0, sigma=0.25)
    x2, v2, _ = r.BumpOnTail.draw_accept_reject_device(20_000, prm, seed=2 ** 40 + 17, ctx=ctx, device_out=False)
    xo2, vo2, _ = smp.sample_bump_on_tail(20_000, seed=2 ** 40 + 17, **prm)
    assert np.array_equal(x2, xo2) and np.max(np.abs(v2 - vo2)) < 1e-13


def test_sampler_chunks_and_device_output(ctx):
    import torch

    n = 50_000
    x, v, w = r.BumpOnTail.draw_accept_reject_device(n, seed=5, ctx=ctx)
    assert x.is_cuda and v.is_cuda and w.is_cuda
    xa, va, wa = r.BumpOnTail.draw_accept_reject_device(10_000, seed=5, ctx=ctx, first=30_000, total=n)
    assert torch.equal(xa, x[30_000:40_000]) and torch.equal(va, v[30_000:40_000]) and torch.equal(wa, w[:10_000])
    # straight into the PIC handle without leaving the device
    L = r.BumpOnTail.domain_length(0.3)
    h = r.PICHandle(n, r.PoissonSolverPBSplines(3, 32, L), ctx)
    h.set_particles(x, v, float(w[0]))
    Wt = np.zeros((1, 11))
    h.run([1.0], 0.1, 10, 11, W=Wt)
    assert np.all(np.isfinite(Wt)) and Wt[0, 0] > 0
    with pytest.raises(r.RBMError):
        r.BumpOnTail.draw_accept_reject_device(10, dict(eps=1.5), ctx=ctx)


def test_growth_rate_matches_oracle(ctx):
    import torch

    r.set_default_context(ctx)
    n, nh, nt, nparam = 20_000, 32, 250, 4
    L = r.BumpOnTail.domain_length(0.3)
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=3)
    particles = r.ParticleList(x[None], v[None], w[None]); poisson = r.PoissonSolverPBSplines(3, nh, L)
    pspace = r.ParameterSpace(r.Parameter("chi", 0.9, 1.2, nparam))
    ip = r.IntegratorParameters(0.1, nt, nt + 1, nh, n, nparam)
    ss = r.integrate_vp_all(particles, poisson, pspace.column("chi"), ip, device="cuda")
    Wh = ss.W.cpu().numpy()                                                   # (nparam, ns)
    nmax = min(len(reg.find_maxima(Wh[i], 2)) for i in range(nparam))
    assert nmax >= 3
    inds = list(range(1, min(nmax, 5)))
    a0, b0 = reg.regression_alpha_beta(ip.t, Wh.T, 2, inds)
    a1, b1 = r.get_regression_alpha_beta_device(ip.t, ss.W, 2, inds, ctx=ctx)      # device traces, no copy
    a2, b2 = r.get_regression_alpha_beta_device(ip.t, Wh, 2, inds, ctx=ctx)        # host traces
    a3, b3 = r.get_regression_alpha_beta(ip.t, Wh.T, 2, inds)                      # host mirror of the reference
    for a, b in ((a1, b1), (a2, b2), (a3, b3)):
        assert np.max(np.abs(a - a0)) < 1e-12 * max(1.0, np.max(np.abs(a0)))
        assert np.max(np.abs(b - b0)) < 1e-12 * max(1.0, np.max(np.abs(b0)))
    with pytest.raises(IndexError):                                              # BoundsError of the reference
        r.get_regression_alpha_beta_device(ip.t, ss.W, 2, [10_000], ctx=ctx)
    # synthetic traces with known growth rates
    t = np.linspace(0.0, 25.0, 251)
    W = np.stack([np.exp(0.3 * t) * (1.2 + np.cos(2.0 * t)), np.exp(0.1 * t) * (1.5 + np.cos(1.5 * t + 0.3))])
    a, b = r.get_regression_alpha_beta_device(t, torch.from_numpy(W).cuda(), 3, [1, 2, 3, 4], ctx=ctx)
    ar, br = reg.regression_alpha_beta(t, W.T, 3, [1, 2, 3, 4])
    assert np.max(np.abs(a - ar)) < 1e-12 and np.max(np.abs(b - br)) < 1e-12
