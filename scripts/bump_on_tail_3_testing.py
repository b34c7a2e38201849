#!/usr/bin/env python
"""Python rendering of scripts/bump_on_tail_3_testing.jl: full model on random test parameters (:45-97), then the
reduced model with DEIM (:108-138), errors between the two."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reducedbasismethods.jl_b200"))

import numpy as np  # noqa: E402

import rbm_b200 as r  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--run", default=os.path.join(ROOT, "runs", "BoT_Np5e4_k_010_050_np_10_T25.npz"))
    ap.add_argument("--ntest", type=int, default=10)
    a = ap.parse_args()
    # everything stage 3 needs comes from the projections file written by stage 2, as in the reference (:30-42):
    # parameter space, parameters, integrator, poisson, the reduced basis and the initial conditions
    ts = rb = r.ReducedBasis.load(a.run.replace(".npz", "_projections.npz"))
    kappa = ts.parameters["kappa"]
    rng = np.random.default_rng(1234)                                         # Random.seed!(1234) (:45)
    chis = np.sort(rng.uniform(0.1, 0.5, a.ntest)) / kappa                    # (:47-57)
    IP = r.IntegratorParameters(ts.integrator.dt, ts.integrator.nt, ts.integrator.ns, ts.integrator.nh,
                                ts.integrator.np, a.ntest)
    t0 = time.perf_counter()
    full = r.integrate_vp_all(ts.initconds, ts.poisson, chis, IP)             # full-order loop (:75-97)
    t_full = time.perf_counter() - t0
    rfield = r.DEIMElectricField(ts.poisson, rb.Psi_p, rb.Psi_e, rb.Pi_e)     # (:132)
    t0 = time.perf_counter()
    red = r.reduced_integrate_vp_all(ts.initconds, rfield, chis, IP)          # (:135)
    t_red = time.perf_counter() - t0
    ex = np.linalg.norm(full.X - red.X) / np.linalg.norm(full.X)
    ew = np.mean(np.abs(full.W - red.W) / np.maximum(full.W, 1e-300))
    print(f" k_p = {rb.k_p}, k_e = {rb.k_e}\n full model  {t_full:.3f} s\n reduced+DEIM {t_red:.3f} s\n"
          f" rel. error X {ex:.3e}, mean rel. error W {ew:.3e}")


if __name__ == "__main__":
    main()
