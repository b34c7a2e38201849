"""Small driver for ncu captures of the reduced-model step (fused reconstruct + deposit kernel) at the configs[4] shape.
usage: RBM_GRAPH=0 python tools/ncu_reduced.py [nt]"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np
import rbm_b200 as r
ctx = r.Context(0); r.set_default_context(ctx)
L = r.BumpOnTail.domain_length(0.3)
n, nh, nt = 100_000, 16, int(sys.argv[1]) if len(sys.argv) > 1 else 4
kp, ke, S = 173, 428, 256
rng = np.random.default_rng(0)
x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=1234)
particles = r.ParticleList(x[None], v[None], w[None]); poisson = r.PoissonSolverPBSplines(3, nh, L)
Psi_p, _ = np.linalg.qr(rng.standard_normal((n, kp))); Psi_e, _ = np.linalg.qr(rng.standard_normal((n, ke)))
idx = rng.choice(n, ke, replace=False)
rfield = r.DEIMElectricField(poisson, np.asfortranarray(Psi_p), np.asfortranarray(Psi_e), idx)
model = r.ReducedModel(particles, rfield, ctx)
chis = np.linspace(0.5, 1.5, S); W = np.zeros((S, 2))
model.run(chis, 0.1, nt, 2, W=W)
print("W_final", W[0, -1])
