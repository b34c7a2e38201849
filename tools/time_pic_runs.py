import os
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np, torch
import rbm_b200 as r
n, nh, nt = 10_000_000, 64, 250
L = r.BumpOnTail.domain_length(0.3)
x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=1234)
ctx = r.Context(0)
h = r.PICHandle(n, r.PoissonSolverPBSplines(3, nh, L), ctx)
xd = torch.from_numpy(x).cuda(); vd = torch.from_numpy(v).cuda()
h.set_particles(xd, vd, float(w[0]))
Xd = torch.empty((1, 2, n), dtype=torch.float64, device="cuda")
Wd = torch.empty((1, 2), dtype=torch.float64, device="cuda")
ts = []
for rep in range(14):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h.run([1.0], 0.1, nt, 2, X=Xd, W=Wd)
    torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print(os.environ.get("RBM_GRAPH", "1"), os.environ.get("RBM_PDL", "1"), " ".join(f"{t:.2f}" for t in ts), "W", Wd.cpu().numpy()[0, -1], "launches", ctx.launch_count())
