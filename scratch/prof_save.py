import os, sys, time
sys.path.insert(0, "reducedbasismethods.jl_b200")
import numpy as np, torch
import rbm_b200 as r
n, nh, nt = 10_000_000, 64, 50
L = r.BumpOnTail.domain_length(0.3)
x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=1234)
ctx = r.Context(0)
h = r.PICHandle(n, r.PoissonSolverPBSplines(3, nh, L), ctx)
h.set_particles(torch.from_numpy(x).cuda(), torch.from_numpy(v).cuda(), float(w[0]))
ns = nt + 1
X = torch.empty((1, ns, n), dtype=torch.float64, device="cuda"); V = torch.empty_like(X); A = torch.empty_like(X)
W = torch.empty((1, ns), dtype=torch.float64, device="cuda")
h.run([1.0], 0.1, nt, ns, X=X, V=V, A=A, W=W)
ctx.profile_enable(True)
torch.cuda.synchronize(); t0 = time.perf_counter()
h.run([1.0], 0.1, nt, ns, X=X, V=V, A=A, W=W)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"profiled run {dt*1e3:.2f} ms for {nt} steps = {dt/nt*1e6:.1f} us/step")
for k in ("k_step", "k_step_save", "k_deposit", "k_solve"):
    c, ms = ctx.profile_query(k); print(f"  {k}: {c} launches, {ms:.2f} ms, {ms/max(c,1)*1e3:.1f} us avg")
