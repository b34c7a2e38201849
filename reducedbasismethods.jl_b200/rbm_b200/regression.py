"""Growth-rate regression (src/regression.jl:4-24) -- host post-processing of W(t)."""
from __future__ import annotations

import numpy as np


def find_maxima(w, n):
    """Indices i (0-based) with w[i-n..i] non-decreasing and w[i..i+n] non-increasing (regression.jl:4-12)."""
    out = []
    for i in range(n, len(w) - n):
        left, right = w[i - n:i + 1], w[i:i + n + 1]
        if np.all(np.diff(left) >= 0) and np.all(np.diff(right) <= 0):
            out.append(i)
    return out


def get_regression_alpha_beta(t, W, n, inds):
    """Least-squares line through log W at selected local maxima; W is (ns, nparam) (regression.jl:14-24)."""
    W = np.asarray(W)
    alpha = np.zeros(W.shape[1]); beta = np.zeros(W.shape[1])
    for i in range(W.shape[1]):
        m = np.asarray(find_maxima(W[:, i], n))[inds]
        lw = np.log(W[m, i])
        beta[i] = np.mean((t[m] - t[m].mean()) * (lw - lw.mean())) / np.var(t[m])
        alpha[i] = lw.mean() - beta[i] * t[m].mean()
    return alpha, beta
