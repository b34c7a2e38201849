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
xh = torch.from_numpy(x).pin_memory(); vh = torch.from_numpy(v).pin_memory()
Xh = torch.empty((1, 2, n), dtype=torch.float64).pin_memory(); Vh = torch.empty_like(Xh).pin_memory(); Ah = torch.empty_like(Xh).pin_memory()
Wh = torch.empty((1, 2), dtype=torch.float64).pin_memory()
ts = []; t_set = []
for rep in range(10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h.set_particles(xh, vh, float(w[0]))
    t1 = time.perf_counter()
    h.run([1.0], 0.1, nt, 2, X=Xh, V=Vh, A=Ah, W=Wh)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    ts.append((t2 - t0) * 1e3); t_set.append((t1 - t0) * 1e3)
print(os.environ.get("RBM_GRAPH", "1"), os.environ.get("RBM_PDL", "1"), "total:", " ".join(f"{t:.1f}" for t in ts), "| set:", " ".join(f"{t:.1f}" for t in t_set))
