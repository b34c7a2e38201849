"""Per-kernel breakdown of one reduced-model step (serialised profile pass) and the graph-replayed step time."""
import os
import sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np, torch
import rbm_b200 as r
ctx = r.Context(0); r.set_default_context(ctx)
L = r.BumpOnTail.domain_length(0.3)
n, nh, nt = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000, 16, 50
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
ts = []
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    model.run(chis, 0.1, nt, 2, W=W)
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print(f"np={n} S={S} kp={kp} ke={ke}: graph-replayed {min(ts)*1e3:.1f} ms for {nt} steps = {min(ts)/nt*1e6:.0f} us/step (RBM_RM_FUSED={os.environ.get('RBM_RM_FUSED','1')}, RBM_RM_FUSE_RECON={os.environ.get('RBM_RM_FUSE_RECON','1')})")
ctx.profile_enable(True)
model.run(chis, 0.1, nt, 2, W=W)
for k in ("k_recon_deposit", "k_dgemm", "k_deposit", "k_solve", "k_rm_field", "k_rm_accel"):
    c, ms = ctx.profile_query(k); print(f"  {k}: {c} launches, {ms:.2f} ms total, {ms/max(c,1)*1e3:.1f} us avg")
ctx.profile_enable(False)
