#!/usr/bin/env python
"""Python rendering of the reference driver scripts/bump_on_tail_1_training.jl on the B200 library:
parameter space -> initial draw -> full-order PIC for every sample -> TrainingSet -> file -> growth rates.
Constants are the shipped defaults (:26-57).  Plots are omitted (no plotting stack in the image)."""
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
    ap.add_argument("--np", type=int, default=int(1e4))      # nb. of particles          (:29)
    ap.add_argument("--nh", type=int, default=16)            # nb. of elements           (:30)
    ap.add_argument("--nparam", type=int, default=10)        # samples of chi            (:44)
    ap.add_argument("--out", default=os.path.join(ROOT, "runs", "BoT_Np5e4_k_010_050_np_10_T25.npz"))
    a = ap.parse_args()
    dt, T, p, kappa = 1e-1, 25, 3, 0.3                       # (:26-31,39)
    nt = int(T // dt)
    params = {"kappa": kappa}
    chi = r.Parameter("chi", 0.1 / kappa, 0.5 / kappa, a.nparam)
    pspace = r.ParameterSpace(chi, r.Parameter("eps", 0.03, 0.03, 1), r.Parameter("a", 0.1, 0.1, 1),
                              r.Parameter("v0", 4.5, 4.5, 1), r.Parameter("sigma", 0.5, 0.5, 1))
    L = 2 * np.pi / kappa                                    # (:64)
    IP = r.VPIntegratorParameters(dt, nt, nt + 1, a.nh, a.np)
    poisson = r.PoissonSolverPBSplines(p, a.nh, L)
    x, v, w = r.BumpOnTail.draw_accept_reject(a.np, seed=1234)          # Random.seed!(1234) (:76-79)
    particles = r.ParticleList(x[None], v[None], w[None])
    TS = r.TrainingSet.create(particles, poisson, nt + 1, params, pspace, IP.with_paramspace(pspace))
    t0 = time.perf_counter()
    r.default_context()                                      # CUDA context + library load: a one-off of the process
    t_init = time.perf_counter() - t0
    t0 = time.perf_counter()
    r.integrate_vp_all(particles, poisson, pspace.column("chi"), TS.integrator, TS.snapshots)   # loop :87-109
    el = time.perf_counter() - t0
    print(f"CUDA context and library: {t_init:.3f} s (once per process)")
    print(f"integrated {len(pspace)} samples x {nt} steps x {a.np} particles in {el:.3f} s "
          f"({len(pspace) * nt * a.np / el:.3e} particle-steps/s), snapshots copied to host arrays")
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    TS.save(a.out)                                           # h5save(file, TS) (:112-114)
    W = TS.snapshots.W.T                                     # (ns, nparam) like the Julia array
    try:
        alpha, beta = r.get_regression_alpha_beta(TS.integrator.t, W, 2, slice(1, 5))   # (:123)
        print("growth rates beta:", np.array2string(beta, precision=4))
    except Exception as e:                                   # too few maxima for tiny runs
        print("regression skipped:", e)


if __name__ == "__main__":
    main()
