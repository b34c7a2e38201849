"""Growth-rate regression (src/regression.jl:4-24): host version and the device entry point."""
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


def get_regression_alpha_beta_device(t, W, n, inds, ctx=None):
    """Same contract on the device (`rbm_growth_rate`): W may be a (nparam, ns) torch CUDA tensor / NumPy array as
    `Snapshots.W` stores it (== Julia's column-major (ns, nparam)) -- nothing is copied for device traces.
    inds: 0-based positions in each trace's list of maxima.  Raises IndexError where the reference throws
    BoundsError."""
    import ctypes as C

    from . import _lib
    from ._lib import RBMError
    from .engine import default_context

    ctx = ctx or default_context()
    if hasattr(W, "data_ptr"):
        nparam, ns = W.shape
        if W.stride(1) != 1:
            raise ValueError("device traces must be stored (nparam, ns) with unit stride along ns")
        ldw, Wp = int(W.stride(0)) if nparam > 1 else int(ns), W
    else:
        Wp = np.ascontiguousarray(W, dtype=np.float64)
        nparam, ns = Wp.shape
        ldw = ns
    tt = np.ascontiguousarray(t, dtype=np.float64)
    if tt.shape[0] != ns:
        raise ValueError("DimensionMismatch: length(t) != size(W, 1)")
    ii = np.ascontiguousarray(np.atleast_1d(inds), dtype=np.int32)
    alpha = np.empty(nparam); beta = np.empty(nparam)
    try:
        ctx.check(ctx.lib.rbm_growth_rate(ctx.h, tt.ctypes.data, _lib.ptr(Wp), ns, nparam, ldw, int(n), ii.ctypes.data,
                                          ii.shape[0], alpha.ctypes.data, beta.ctypes.data, None))
    except RBMError as e:
        if "BoundsError" in str(e):
            raise IndexError(str(e)) from None
        raise
    return alpha, beta
