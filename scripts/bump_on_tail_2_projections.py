#!/usr/bin/env python
"""Python rendering of scripts/bump_on_tail_2_projections.jl: TrainingSet -> ReducedBasis(CotangentLiftEVD) ->
file -> projected snapshots (:22-60)."""
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
    ap.add_argument("--particle-tol", type=float, default=1e-4)       # (:22)
    ap.add_argument("--field-tol", type=float, default=1e-2)          # (:23)
    a = ap.parse_args()
    ts = r.TrainingSet.load(a.run)                                    # (:25)
    t0 = time.perf_counter()
    rb = r.ReducedBasis.from_trainingset(r.CotangentLiftEVD(), ts, particle_tol=a.particle_tol, field_tol=a.field_tol)
    print(f" k_p = {rb.k_p}\n k_e = {rb.k_e}\n ({time.perf_counter() - t0:.3f} s)")       # (:28-30)
    rb.save(a.run.replace(".npz", "_projections.npz"))                # (:40)
    Xt = r.project(rb.Psi_p, ts.snapshots.matrix("X"))                # X~ = Psi_p' X  (:44-60)
    At = r.project(rb.Psi_p, ts.snapshots.matrix("A"))
    np.savez(a.run.replace(".npz", "_vectorfield.npz"), X=Xt, A=At)


if __name__ == "__main__":
    main()
