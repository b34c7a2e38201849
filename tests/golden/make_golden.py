"""Generates the committed golden fixtures under tests/golden/.

There is no runnable reference here (no Julia; the PIC arithmetic is in un-vendored registry
packages), so the fixtures are produced by the NumPy restatements under oracle/ and pin (a) the C
oracle against an independent implementation and (b) the GPU path against fixed numbers that do not
depend on anything being rebuilt.  The POD/DEIM fixture uses the reference's own known-answer input
(test/deim_test.jl:4-12).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "reducedbasismethods.jl_b200"))

from oracle import pic_numpy as pn, pod_numpy as pod          # noqa: E402
from rbm_b200.bump_on_tail import domain_length, draw_accept_reject  # noqa: E402


def pic_fixture():
    L = domain_length(0.3)
    n, nh, nt, ns, dt = 2000, 16, 40, 11, 0.1
    x, v, w = draw_accept_reject(n, seed=1234)
    out = {"x0": x, "v0": v, "w": w[0], "L": L, "nh": nh, "nt": nt, "ns": ns, "dt": dt}
    chis = np.array([1.0 / 3.0, 1.0, 5.0 / 3.0])
    out["chi"] = chis
    for p, chi in enumerate(chis):
        o = pn.integrate(x, v, w, L, nh, float(chi), dt, nt, ns)
        for k in ("W", "K", "M", "Phi"):
            out[f"{k}_{p}"] = o[k]
        out[f"Xlast_{p}"] = o["X"][-1]
        out[f"Vlast_{p}"] = o["V"][-1]
        out[f"Alast_{p}"] = o["A"][-1]
    # single-kernel fixtures
    xs = np.concatenate([x[:500], -x[:200], 7.3 * x[:200],
                         np.array([0.0, -0.0, L, -L, 2 * L, L / nh, 3 * L / nh, -1e-300, 1e-300,
                                   np.nextafter(L, 0), np.nextafter(L, 2 * L), -np.nextafter(L, 0)])])
    c, xi = pn.locate(xs, L, nh)
    out["loc_x"] = xs; out["loc_cell"] = c; out["loc_xi"] = xi
    rhs = pn.deposit(x, w, L, nh)
    phi, W = pn.poisson(rhs, nh, L, 0.75)
    out["dep_rhs"] = rhs; out["sol_phi"] = phi; out["sol_W"] = W
    out["gat_a"] = pn.gather(phi, x, L, nh)
    np.savez_compressed(os.path.join(HERE, "pic_small.npz"), **out)


def pod_fixture():
    # reference KAT input: test/deim_test.jl:4-12
    xg = np.linspace(-1, 1, 100)
    mu = np.linspace(1, np.pi, 51)
    S = (1 - xg[:, None]) * np.cos(3 * np.pi * mu[None, :] * (xg[:, None] + 1)) * np.exp(-(1 + xg[:, None]) * mu[None, :])
    Psi, Lam = pod.pod_evd(S, k=20)
    idx = pod.deim_indices(Psi)
    np.savez_compressed(os.path.join(HERE, "pod_deim_kat.npz"), S=S, Psi=Psi, Lam=Lam, idx=idx)


if __name__ == "__main__":
    pic_fixture()
    pod_fixture()
    print("wrote", os.listdir(HERE))
