import sys, time
sys.path.insert(0, "reducedbasismethods.jl_b200")
import numpy as np, torch
import rbm_b200 as r
ctx = r.Context(0)
rows, cols = 1_000_000, 2000
m = cols // 2
X = torch.randn(m, rows, dtype=torch.float64, device="cuda").T
V = torch.randn(m, rows, dtype=torch.float64, device="cuda").T
G = torch.empty(cols, cols, dtype=torch.float64, device="cuda")
keep, P, N, Ld, nb, nrows, C = r.reducedbasis._blocks((X, V))
for rep in range(5):
    if rep == 3: ctx.profile_enable(True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.check(ctx.lib.rbm_gram(ctx.h, P, N, Ld, nb, nrows, G.data_ptr()))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(rep, f"{dt*1e3:.2f} ms", torch.cuda.mem_get_info())
