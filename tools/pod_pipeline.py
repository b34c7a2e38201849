import os
import sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np, torch
import rbm_b200 as r
ctx = r.Context(0); r.set_default_context(ctx)
L = r.BumpOnTail.domain_length(0.3)
def T(f, *a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = f(*a, **k); torch.cuda.synchronize(); return out, time.perf_counter() - t0
for n in (5000, 10_000, 100_000):
    nh, nt, nparam = 16, 250, 10
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=1234)
    particles = r.ParticleList(x[None], v[None], w[None]); poisson = r.PoissonSolverPBSplines(3, nh, L)
    pspace = r.ParameterSpace(r.Parameter("chi", 1/3, 5/3, nparam))
    ip = r.IntegratorParameters(0.1, nt, nt + 1, nh, n, nparam)
    ts = r.TrainingSet.create(particles, poisson, nt + 1, {"kappa": 0.3}, pspace, ip)
    _, t_pic = T(r.integrate_vp_all, particles, poisson, pspace.column("chi"), ip, ts.snapshots)
    X, V, E = ts.snapshots.matrix("X"), ts.snapshots.matrix("V"), ts.snapshots.matrix("A")
    (Pp, Lp), t_p = T(r.get_PODBasis_cotangentLiftEVD, X, V, tolerance=1e-9)
    (Pe, Le), t_e = T(r.get_PODBasis_EVD, E, tolerance=1e-5)
    idx, t_d = T(r.deim_indices, Pe)
    print(f"np={n}: PIC {t_pic:.3f}s | cotangent-lift POD (C={2*X.shape[1]}) {t_p:.3f}s k={Pp.shape[1]} | field POD {t_e:.3f}s k={Pe.shape[1]} | DEIM {t_d:.3f}s")
