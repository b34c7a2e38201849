"""Fixed per-launch cost of the step kernel: time per step against the number of 1536-particle tiles per block (np =
148 * 1536 * tiles), graph replay.  The intercept is what every PIC step pays regardless of the particle count."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np, torch
import rbm_b200 as r
L = r.BumpOnTail.domain_length(0.3)
ctx = r.Context(0)
nt = 1000
for tiles in ([int(t) for t in sys.argv[1].split(",")] if len(sys.argv) > 1 else (1, 2, 4, 8, 16, 44)):
    n = 148 * 1536 * tiles
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=1)
    h = r.PICHandle(n, r.PoissonSolverPBSplines(3, int(os.environ.get("NH", "64")), L), ctx)
    h.set_particles(torch.from_numpy(x).cuda(), torch.from_numpy(v).cuda(), float(w[0]))
    W = torch.empty((1, 2), dtype=torch.float64, device="cuda")
    best = 1e9
    for rep in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        h.run([1.0], 0.1, nt, 2, W=W)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    print(f"tiles/block={tiles:3d} np={n:9d}: {best/nt*1e6:7.2f} us per step")
    h.close()
