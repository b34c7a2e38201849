"""TEST INFRASTRUCTURE ONLY (oracle): literal restatement of /root/reference/src/regression.jl:4-24
(find_maxima, get_regression_alpha_beta) -- in-tree reference code, so this half is pinned by reading it."""
from __future__ import annotations

import numpy as np


def find_maxima(w, n):
    """regression.jl:4-12; 0-based indices.  issorted(a, lt=isless): no i with a[i+1] < a[i]."""
    out = []
    for i in range(n, len(w) - n):
        left, right = w[i - n:i + 1], w[i:i + n + 1]
        if not np.any(left[1:] < left[:-1]) and not np.any(right[:-1] < right[1:]):
            out.append(i)
    return out


def regression_alpha_beta(t, W, n, inds):
    """regression.jl:14-24; W is (ns, nparam); inds 0-based positions in the maxima list."""
    W = np.asarray(W, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    alpha = np.zeros(W.shape[1])
    beta = np.zeros(W.shape[1])
    for i in range(W.shape[1]):
        m = np.asarray(find_maxima(W[:, i], n))[np.asarray(inds)]      # IndexError == Julia BoundsError
        lw = np.log(W[m, i])
        cov = np.mean((t[m] - t[m].mean()) * (lw - lw.mean()))         # cov(..., corrected=false)
        var = np.mean((t[m] - t[m].mean()) ** 2)                       # var(..., corrected=false)
        beta[i] = cov / var
        alpha[i] = lw.mean() - beta[i] * t[m].mean()
    return alpha, beta
