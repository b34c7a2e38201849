import os
import sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np, torch
import rbm_b200 as r
L = r.BumpOnTail.domain_length(0.3)
ctx = r.Context(0)
def run(n, nh, nt, ns, weighted=False, outputs="XVA"):
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=1)
    h = r.PICHandle(n, r.PoissonSolverPBSplines(3, nh, L), ctx)
    wd = torch.from_numpy(w).cuda() if weighted else float(w[0])
    h.set_particles(torch.from_numpy(x).cuda(), torch.from_numpy(v).cuda(), wd)
    out = {k: torch.empty((1, ns, n), dtype=torch.float64, device="cuda") for k in outputs}
    Wd = torch.empty((1, ns), dtype=torch.float64, device="cuda")
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        h.run([1.0], 0.1, nt, ns, W=Wd, **out)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    print(f"np={n} nh={nh} nt={nt} ns={ns} weighted={weighted} outputs={outputs!r}: {best*1e3:.2f} ms, {best/nt*1e6:.1f} us/step, {nt*n/best:.3e} particle-steps/s")
    h.close(); del out
run(10_000_000, 64, 250, 2)
run(10_000_000, 64, 250, 2, weighted=True)
run(10_000_000, 16, 250, 2)
run(10_000_000, 128, 100, 2)
run(10_000_000, 256, 100, 2)
run(10_000_000, 64, 100, 101)            # every step saved, X V A resident on the device (24 GB)
run(10_000_000, 64, 100, 11)
run(10_000_000, 100, 100, 2)
run(10_000_000, 128, 100, 101)
