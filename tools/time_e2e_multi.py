"""Training-set generation end to end with host snapshot arrays: S samples x np = 1e7, ns = 2, pinned X, V, A.
Run with RBM_STAGE_SETS2=1 (two alternating staging sets, default) and =0 (one set, copy stream drained per group)."""
import os
import sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np, torch
import rbm_b200 as r
S = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n, nh, nt, ns = 10_000_000, 64, 250, 2
L = r.BumpOnTail.domain_length(0.3)
x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=1234)
ctx = r.Context(0)
h = r.PICHandle(n, r.PoissonSolverPBSplines(3, nh, L), ctx)
xh = torch.from_numpy(x).pin_memory(); vh = torch.from_numpy(v).pin_memory()
Xh = torch.empty((S, ns, n), dtype=torch.float64).pin_memory(); Vh = torch.empty_like(Xh).pin_memory(); Ah = torch.empty_like(Xh).pin_memory()
Wh = torch.empty((S, ns), dtype=torch.float64).pin_memory()
chi = np.linspace(1 / 3, 5 / 3, S)
for sets in ("1", "0", "1", "0"):
    os.environ["RBM_STAGE_SETS2"] = sets
    ts = []
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        h.set_particles(xh, vh, float(w[0]))
        h.run(chi, 0.1, nt, ns, X=Xh, V=Vh, A=Ah, W=Wh)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    best = min(ts[1:])
    print(f"S={S} staging sets={'2' if sets == '1' else '1'}: " + " ".join(f"{t:.1f}" for t in ts) + f" ms | {S * nt * n / best / 1e-3:.3e} particle-steps/s end to end", flush=True)
