"""CPU tests of the sampler / regression oracles (test infrastructure): Philox4x32-10 against the Random123
known-answer vectors, the sampled distribution against its analytic moments, and the regression restatement
against a hand-checkable trace."""
import numpy as np

from oracle import regression_numpy as reg
from oracle import sampler_numpy as smp


def test_philox_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        r = smp.philox4x32_10(np.array([ctr], dtype=np.uint32), key)[0]
        assert tuple(int(v) for v in r) == out


def test_sampled_distribution_moments():
    n = 200_000
    x, v, rej = smp.sample_bump_on_tail(n, seed=7)
    L = 2 * np.pi / 0.3
    assert x.min() >= 0.0 and x.max() < L
    assert abs(v.mean() - 0.45) < 4 * np.sqrt(2.7475 / n)
    assert abs(v.var() - (0.9 + 0.1 * (0.25 + 4.5 ** 2) - 0.45 ** 2)) < 0.05
    assert abs(np.cos(0.3 * x).mean() - 0.015) < 4 / np.sqrt(2 * n)          # E cos(kx) = eps/2
    assert abs(rej / (n + rej) - 0.03 / 1.03) < 2e-3                         # rejection rate eps/(1+eps)
    # chunks of one global sample: particle p depends only on (seed, p)
    xa, va, _ = smp.sample_bump_on_tail(1000, seed=7, first=5000)
    assert np.array_equal(xa, x[5000:6000]) and np.array_equal(va, v[5000:6000])


def test_regression_restatement():
    t = np.linspace(0.0, 25.0, 251)
    W = np.stack([np.exp(0.3 * t) * (1.2 + np.cos(2.0 * t)), np.exp(0.1 * t) * (1.5 + np.cos(1.5 * t + 0.3))], axis=1)
    m = reg.find_maxima(W[:, 0], 3)
    assert len(m) >= 5 and all(W[i, 0] >= W[i - 1, 0] and W[i, 0] >= W[i + 1, 0] for i in m)
    alpha, beta = reg.regression_alpha_beta(t, W, 3, [1, 2, 3, 4])
    assert abs(beta[0] - 0.3) < 0.02 and abs(beta[1] - 0.1) < 0.02
    # plateau: issorted accepts ties on both sides -> every point of a flat top is a "maximum" (reference semantics)
    w = np.array([0.0, 1.0, 2.0, 2.0, 1.0, 0.0])
    assert reg.find_maxima(w, 1) == [2, 3]
