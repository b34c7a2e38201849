"""oracle/pod_numpy.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy restatement of the reference's snapshot POD (Gram + EVD) and DEIM point
selection.  This half of the path IS in /root/reference and is pinned by the
reference's own known-answer test test/deim_test.jl:1-36 (restated in
tests/test_oracle_pod.py: eigenvalue-decay bounds :14-18 and the DEIM mean
interpolation error < 5e-5 :20-36).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.
"""
from __future__ import annotations

import numpy as np


def sorteigen(evals, evecs):
    """src/eigen.jl:4-7 -- permutation by |lambda| descending (stable)."""
    p = np.argsort(-np.abs(evals), kind="stable")
    return evals[p], evecs[:, p]


def _rank_from_tolerance(Lam, tolerance):
    """src/algorithms/evd.jl:38-46 / :70-78 -- first k with
    sum_{i<=k} Lam_i / sum(Lam) >= 1 - tol (E includes round-off negatives)."""
    E = np.sum(Lam)
    Er = 0.0
    k = 1
    while 1 - tolerance > (Er + Lam[k - 1]) / E:
        Er += Lam[k - 1]
        k += 1
    return k


def gram(*blocks):
    """[B1 B2 ...]' * [B1 B2 ...] -- src/algorithms/evd.jl:32-33 (cotangent lift
    [X V]) and :65 (single matrix)."""
    S = np.hstack(blocks)
    return S.T @ S


def pod_evd(S, k=0, tolerance=1e-4):
    """get_PODBasis_EVD -- src/algorithms/evd.jl:63-87.  Returns (Psi, Lam) with ALL
    eigenvalues in Lam (:86).  `eigen` takes LAPACK's symmetric path for a
    Hermitian Gram; restated with eigh on the symmetrised matrix."""
    G = S.T @ S
    G = 0.5 * (G + G.T)
    vals, vecs = np.linalg.eigh(G)
    Lam, Om = sorteigen(vals, vecs)
    if k == 0:
        k = _rank_from_tolerance(Lam, tolerance)
    Psi = S @ Om[:, :k] @ np.diag(1.0 / np.sqrt(np.abs(Lam[:k])))
    return Psi, Lam


def pod_cotangent_lift_evd(X, V, k=0, tolerance=1e-8):
    """get_PODBasis_cotangentLiftEVD -- src/algorithms/evd.jl:26-55."""
    assert X.shape == V.shape
    return pod_evd(np.hstack([X, V]), k=k, tolerance=tolerance)


def deim_indices(Psi):
    """get_DEIM_interpolation_matrix -- src/algorithms/deim.jl:6-21, returning the
    selected rows (0-based) instead of the dense N x k 0/1 matrix.  NB the
    reference uses the SIGNED argmax (:8,:16)."""
    N, k = Psi.shape
    idx = [int(np.argmax(Psi[:, 0]))]
    for i in range(1, k):
        P = Psi[idx, :i]
        c = np.linalg.solve(P, Psi[idx, i])
        r = Psi[:, i] - Psi[:, :i] @ c
        idx.append(int(np.argmax(r)))
    return np.array(idx, dtype=np.int64)


def deim_matrix(Psi):
    """Dense Pi (N x k) exactly as the reference returns it (deim.jl:9,18)."""
    idx = deim_indices(Psi)
    Pi = np.zeros(Psi.shape)
    Pi[idx, np.arange(Psi.shape[1])] = 1.0
    return Pi


def reduced_basis(X, V, A, particle_tol=1e-8, field_tol=1e-4):
    """ReducedBasis(::CotangentLiftEVD, ts) -- src/algorithms/evd.jl:119-137 on the
    already-reshaped (np, ns*nparam) matrices (:124-126)."""
    Psi_p, Lam_p = pod_cotangent_lift_evd(X, V, tolerance=particle_tol)
    Psi_e, Lam_e = pod_evd(A, tolerance=field_tol)
    Pi_e = deim_indices(Psi_e)
    return dict(Psi_p=Psi_p, Lam_p=Lam_p, Psi_e=Psi_e, Lam_e=Lam_e, idx_e=Pi_e)
