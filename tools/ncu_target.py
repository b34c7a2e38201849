"""Small driver for ncu captures of the PIC kernels at the bench size (np = 1e7, nh = 64).
usage: python tools/ncu_target.py step|dual [nt]      (RBM_GRAPH=0 RBM_PDL=0 make every launch individually visible)"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np, torch
import rbm_b200 as r
mode = sys.argv[1] if len(sys.argv) > 1 else "step"
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n, nh = 10_000_000, 64
L = r.BumpOnTail.domain_length(0.3)
x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=1234)
ctx = r.Context(0)
h = r.PICHandle(n, r.PoissonSolverPBSplines(3, nh, L), ctx)
h.set_particles(torch.from_numpy(x).cuda(), torch.from_numpy(v).cuda(), float(w[0]))
ns = nt + 1 if mode == "dual" else 2
X = torch.empty((1, ns, n), dtype=torch.float64, device="cuda"); V = torch.empty_like(X); A = torch.empty_like(X)
W = torch.empty((1, ns), dtype=torch.float64, device="cuda")
h.run([1.0], 0.1, nt, ns, X=X, V=V, A=A, W=W)
torch.cuda.synchronize()
print(mode, "W_final", float(W[0, -1]))
