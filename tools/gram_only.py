import os
import sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np, torch
import rbm_b200 as r
ctx = r.Context(0)
for rows, cols in [(200_000, 2000), (1_000_000, 2000), (1_000_000, 502), (2_000_000, 1000), (100_000, 5020)]:
    m = cols // 2
    X = torch.randn(m, rows, dtype=torch.float64, device="cuda").T
    V = torch.randn(m, rows, dtype=torch.float64, device="cuda").T
    G = torch.empty(cols, cols, dtype=torch.float64, device="cuda")
    keep, P, N, Ld, nb, nrows, C = r.reducedbasis._blocks((X, V))
    ctx.check(ctx.lib.rbm_gram(ctx.h, P, N, Ld, nb, nrows, G.data_ptr()))
    ctx.profile_enable(True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.check(ctx.lib.rbm_gram(ctx.h, P, N, Ld, nb, nrows, G.data_ptr()))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    n, ms = ctx.profile_query("k_dgemm"); ctx.profile_enable(False)
    fl = rows*cols*(cols+1)
    print(f"gram {rows}x{cols}: total {dt*1e3:.2f} ms {fl/dt/1e12:.2f} TF | kernel {ms:.2f} ms {fl/ms/1e9:.2f} TF")
    del X, V, G
