"""CPU: the NumPy POD/DEIM oracle against the reference's own known-answer test
(test/deim_test.jl:1-36) and the committed golden vectors."""
import os

import numpy as np

from oracle import pod_numpy as pod


def _s(x, mu):
    return (1 - x) * np.cos(3 * np.pi * mu * (x + 1)) * np.exp(-(1 + x) * mu)


def test_reference_deim_kat():
    x = np.linspace(-1, 1, 100)
    mu_t = np.linspace(1, np.pi, 51)
    mu_v = np.linspace(1, np.pi, 101)
    S = _s(x[:, None], mu_t[None, :])
    Psi, Lam = pod.pod_evd(S, k=20)
    assert Lam[9] < 1e0 and Lam[14] < 1e-2 and Lam[19] < 1e-7        # deim_test.jl:14-18
    Pi = pod.deim_matrix(Psi)
    Sv = _s(x[:, None], mu_v[None, :])
    D = Psi @ np.linalg.inv(Pi.T @ Psi)
    xe = Pi.T @ x
    Sa = np.stack([D @ _s(xe, m) for m in mu_v], axis=1)
    err = np.mean(np.linalg.norm(Sv - Sa, axis=0))
    assert err < 5e-5                                                 # deim_test.jl:34-36


def test_golden_pod(golden_dir):
    g = np.load(os.path.join(golden_dir, "pod_deim_kat.npz"))
    Psi, Lam = pod.pod_evd(g["S"], k=20)
    assert np.allclose(Lam[:12], g["Lam"][:12], rtol=1e-9)
    assert np.array_equal(pod.deim_indices(g["Psi"]), g["idx"])
    assert np.allclose(Psi.T @ Psi, np.eye(20), atol=1e-5)


def test_rank_from_tolerance_and_cotangent_lift():
    rng = np.random.default_rng(0)
    U, _ = np.linalg.qr(rng.standard_normal((300, 12)))
    Wm, _ = np.linalg.qr(rng.standard_normal((40, 12)))
    sig = 10.0 ** -np.arange(12)
    S = U @ np.diag(sig) @ Wm.T
    Psi, Lam = pod.pod_evd(S, tolerance=1e-6)
    assert np.allclose(Lam[:4], sig[:4] ** 2, rtol=1e-8)
    assert Psi.shape[1] == 3          # (1 + 1e-2 + 1e-4)/sum = 1 - 9.9e-7 >= 1 - 1e-6
    Psi2, Lam2 = pod.pod_cotangent_lift_evd(S[:, :20], S[:, 20:], tolerance=1e-6)
    assert Lam2.shape[0] == 40 and Psi2.shape[0] == 300


def test_sorteigen_orders_by_abs_descending():
    vals = np.array([1.0, -3.0, 2.0, -1e-15])
    vecs = np.eye(4)
    l, v = pod.sorteigen(vals, vecs)
    assert list(l) == [-3.0, 2.0, 1.0, -1e-15]
    assert np.array_equal(v, vecs[:, [1, 2, 0, 3]])
