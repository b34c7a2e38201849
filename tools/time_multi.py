import os
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np, torch
import rbm_b200 as r
L = r.BumpOnTail.domain_length(0.3)
ctx = r.Context(0)
for n, nparam, nh, nt, ns in [(10_000_000, 4, 64, 100, 2), (1_000_000, 16, 64, 250, 2), (100_000, 64, 64, 250, 2), (10_000, 10, 16, 250, 251), (10_000, 256, 16, 250, 2), (2_000_000, 8, 64, 100, 11)]:
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=1)
    h = r.PICHandle(n, r.PoissonSolverPBSplines(3, nh, L), ctx)
    h.set_particles(torch.from_numpy(x).cuda(), torch.from_numpy(v).cuda(), float(w[0]))
    chis = np.linspace(1/3, 5/3, nparam)
    Wd = torch.empty((nparam, ns), dtype=torch.float64, device="cuda")
    Xd = torch.empty((nparam, ns, n), dtype=torch.float64, device="cuda") if ns > 2 else None
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        h.run(chis, 0.1, nt, ns, W=Wd, X=Xd)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    print(f"np={n} nparam={nparam} nh={nh} nt={nt} ns={ns}: {best*1e3:.2f} ms  {nparam*nt*n/best:.3e} particle-steps/s")
    h.close(); del Xd
