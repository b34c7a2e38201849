"""Stage 1 -> stage 2 hand-over: snapshots kept in HBM vs the host round trip.
usage: python tools/handover.py [np] [nparam] [ns]"""
import os
import sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np, torch
import rbm_b200 as r
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nparam = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 26
nt, nh = 250, 64
ctx = r.Context(0); r.set_default_context(ctx)
L = r.BumpOnTail.domain_length(0.3)
x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=1)
particles = r.ParticleList(x[None], v[None], w[None]); poisson = r.PoissonSolverPBSplines(3, nh, L)
pspace = r.ParameterSpace(r.Parameter("chi", 1 / 3, 5 / 3, nparam))
ip = r.IntegratorParameters(0.1, nt, ns, nh, n, nparam)
def T(f, *a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = f(*a, **k); torch.cuda.synchronize(); return out, time.perf_counter() - t0
gb = 3 * nparam * ns * n * 8 / 1e9
for dev in ("cuda", None, "cuda", None):
    ts = r.TrainingSet.create(particles, poisson, ns, {"kappa": 0.3}, pspace, ip, device=dev)
    if dev is None:
        for k in ("X", "V", "A"): getattr(ts.snapshots, k).fill(0.0)          # touch the pages (first-touch cost is not ours)
    _, t1 = T(r.integrate_vp_all, particles, poisson, pspace.column("chi"), ip, ts.snapshots)
    rb, t2 = T(r.ReducedBasis.from_trainingset, r.CotangentLiftEVD(), ts)
    print(f"np={n} nparam={nparam} ns={ns} ({gb:.1f} GB of X,V,A) snapshots on {'device' if dev else 'host  '}: "
          f"stage 1 {t1:.3f} s, stage 2 (2 POD + DEIM) {t2:.3f} s, k_p={rb.k_p} k_e={rb.k_e}", flush=True)
    del ts, rb
