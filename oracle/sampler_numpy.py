"""TEST INFRASTRUCTURE ONLY (oracle): NumPy restatement of the device initial-condition sampler
(reducedbasismethods.jl_b200/csrc/sampling.cu, `rbm_sample_bump_on_tail`).

What it stands in for: VlasovMethods.BumpOnTail.draw_accept_reject(np, params) as called at
/root/reference/scripts/bump_on_tail_1_training.jl:76-79.  That function lives in an un-vendored package
and consumes Julia's MersenneTwister stream, which no GPU sampler can reproduce; the distribution
    f(x, v) = (1 + eps cos(kappa x)) [(1-a) N(0,1)(v) + a N(v0, sigma^2)(v)],  x in [0, 2 pi / kappa)
is what the reference fixes (notebooks/VP-ROM-Projections.ipynb cell 2), the stream is ours:
Philox4x32-10 (Salmon et al., SC'11; known-answer vectors of the Random123 distribution are checked in
tests/test_oracle_sampler.py), key = seed, counter = (particle index, draw number).
  draw 0: (u0, _)   tail component iff u0 < a
  draw 1: (u2, u3)  z = sqrt(-2 log(1-u2)) cos(2 pi u3);  v = tail ? v0 + sigma z : z
  draw j >= 2: (uc, ua)  candidate x = uc L, accepted iff ua (1+eps) <= 1 + eps cos(kappa x)
PARITY: unpinned against Julia (see above); pinned as a generator by the Philox known-answer vectors and as a
distribution by moment tests.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32, key: (2,) uint32 -> (..., 4) uint32."""
    c = [np.asarray(ctr[..., i], dtype=np.uint32).copy() for i in range(4)]
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return np.stack(c, axis=-1)


def _uniform_pairs(p, j, seed):
    """Two 53-bit uniforms in [0,1) per particle index p (uint64 array) for draw number j."""
    ctr = np.empty(p.shape + (4,), dtype=np.uint32)
    ctr[..., 0] = (p & MASK).astype(np.uint32)
    ctr[..., 1] = (p >> np.uint64(32)).astype(np.uint32)
    ctr[..., 2] = np.uint32(j)
    ctr[..., 3] = 0
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).astype(np.uint64)
    u0 = (((r[..., 0] << np.uint64(32)) | r[..., 1]) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    u1 = (((r[..., 2] << np.uint64(32)) | r[..., 3]) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    return u0, u1


def sample_bump_on_tail(n, seed=1234, kappa=0.3, eps=0.03, a=0.1, v0=4.5, sigma=0.5, first=0):
    """-> x, v, w (uniform L/n is the caller's business: returned as None), rejected count."""
    L = 2.0 * 3.14159265358979323846 / kappa
    p = np.arange(first, first + n, dtype=np.uint64)
    u0, _ = _uniform_pairs(p, 0, seed)
    u2, u3 = _uniform_pairs(p, 1, seed)
    z = np.sqrt(-2.0 * np.log(1.0 - u2)) * np.cos(6.283185307179586 * u3)
    v = np.where(u0 < a, v0 + sigma * z, z)
    x = np.empty(n, dtype=np.float64)
    todo = np.arange(n)
    rejected, j = 0, 2
    while todo.size:
        uc, ua = _uniform_pairs(p[todo], j, seed)
        cand = uc * L
        ok = ua * (1.0 + eps) <= 1.0 + eps * np.cos(kappa * cand)
        x[todo[ok]] = cand[ok]
        rejected += int((~ok).sum())
        todo = todo[~ok]
        j += 1
    return x, v, rejected
