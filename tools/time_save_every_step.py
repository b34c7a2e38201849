"""BASELINE configs[1], second run of SURVEY 8(d): ns = nt + 1 (every step saved: +24 B per particle-step, 60 GB of
device-resident snapshots at np = 1e7)."""
import os
import sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np, torch
import rbm_b200 as r
n, nh, nt = 10_000_000, 64, 250
L = r.BumpOnTail.domain_length(0.3)
x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=1234)
ctx = r.Context(0)
h = r.PICHandle(n, r.PoissonSolverPBSplines(3, nh, L), ctx)
h.set_particles(torch.from_numpy(x).cuda(), torch.from_numpy(v).cuda(), float(w[0]))
for ns in (2, 26, 251):
    X = torch.empty((1, ns, n), dtype=torch.float64, device="cuda"); V = torch.empty_like(X); A = torch.empty_like(X)
    W = torch.empty((1, ns), dtype=torch.float64, device="cuda")
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        h.run([1.0], 0.1, nt, ns, X=X, V=V, A=A, W=W)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    byt = n * (32.0 * nt + 24.0 * ns)
    print(f"ns={ns}: {best*1e3:.2f} ms, {n*nt/best:.3e} particle-steps/s, algorithmic {byt/best/1e9:.0f} GB/s ({byt/best/1e9/6541.1:.2f} of the measured HBM peak)")
    del X, V, A
