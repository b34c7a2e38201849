"""Small end-to-end exercise of the round-2 kernels for compute-sanitizer (memcheck / racecheck): saved steps with the
second histogram (paired and full), large-grid paired main kernel, blocked DEIM (several panels), fused reduced step."""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np
import rbm_b200 as r
ctx = r.Context(0); r.set_default_context(ctx)
L = r.BumpOnTail.domain_length(0.3)
for n, nh in ((6001, 64), (5000, 16), (4099, 128)):
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=n)
    h = r.PICHandle(n, r.PoissonSolverPBSplines(3, nh, L), ctx)
    h.set_particles(x, v, float(w[0]))
    nt, ns = 6, 7
    X = np.zeros((2, ns, n)); W = np.zeros((2, ns)); Phi = np.zeros((2, ns, nh))
    h.run([0.8, 1.2], 0.1, nt, ns, X=X, W=W, Phi=Phi)
    h.run([0.8, 1.2], 0.1, nt, 3, W=np.zeros((2, 3)))
    print("pic", n, nh, W[:, -1])
    h.close()
rng = np.random.default_rng(0)
Q = np.asfortranarray(np.linalg.qr(rng.standard_normal((1501, 70)))[0])
print("deim", r.deim_indices(Q)[:5])
n = 3000
x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=9)
pr = r.ParticleList(x[None], v[None], w[None]); poi = r.PoissonSolverPBSplines(3, 16, L)
Pp = np.asfortranarray(np.linalg.qr(rng.standard_normal((n, 21)))[0]); Pe = np.asfortranarray(np.linalg.qr(rng.standard_normal((n, 37)))[0])
rf = r.DEIMElectricField(poi, Pp, Pe, r.deim_indices(Pe))
ip = r.IntegratorParameters(0.1, 5, 6, 16, n, 3)
rss = r.reduced_integrate_vp_all(pr, rf, [0.9, 1.0, 1.1], ip)
print("reduced", rss.W[:, -1])
# staged host snapshots with two alternating staging sets (four lock-step groups of one sample)
os.environ["RBM_GROUP"] = "1"
n = 7001
x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=3)
h = r.PICHandle(n, r.PoissonSolverPBSplines(3, 16, L), ctx)
h.set_particles(x, v, float(w[0]))
X = np.zeros((4, 3, n)); V = np.zeros_like(X); A = np.zeros_like(X)
h.run([0.7, 0.9, 1.1, 1.3], 0.1, 4, 3, X=X, V=V, A=A)
print("staged", float(np.abs(X).max()))
h.close()
del os.environ["RBM_GROUP"]
# fused reconstruct + deposit (needs at least two tiles per SM: 157 row tiles x 2 sample tiles)
n, S = 20_000, 128
x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=10)
pr = r.ParticleList(x[None], v[None], w[None])
Pp = np.asfortranarray(np.linalg.qr(rng.standard_normal((n, 19)))[0]); Pe = np.asfortranarray(np.linalg.qr(rng.standard_normal((n, 23)))[0])
rf = r.DEIMElectricField(poi, Pp, Pe, r.deim_indices(Pe))
ip = r.IntegratorParameters(0.1, 3, 2, 16, n, S)
rss = r.reduced_integrate_vp_all(pr, rf, np.linspace(0.5, 1.5, S), ip)
print("reduced fused", rss.W[:3, -1])
