"""DEIM point selection: bytes streamed (sum_i 8 N (i+1)) against the HBM peak, and the per-iteration overhead at small N.
usage: python tools/time_deim.py"""
import os
import sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np, torch
import rbm_b200 as r
ctx = r.Context(0); r.set_default_context(ctx)
for n, k in ((10_000, 724), (100_000, 256), (1_000_000, 128), (10_000_000, 64)):
    torch.manual_seed(0)
    Q, _ = torch.linalg.qr(torch.randn(n, k, dtype=torch.float64, device="cuda"))
    Psi = Q.T.contiguous().T
    r.deim_indices(Psi)
    ts = []
    for rep in range(3):      # best of three: single calls occasionally catch a multi-100-ms hiccup of the shared box
        torch.cuda.synchronize(); t0 = time.perf_counter()
        idx = r.deim_indices(Psi)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    dt = min(ts)
    byt = sum(8.0 * n * (i + 1) + (8.0 * n if i else 0) for i in range(k))
    print(f"N={n} k={k}: {dt*1e3:.2f} ms ({dt/k*1e6:.1f} us per point), streamed {byt/1e9:.2f} GB -> {byt/dt/1e9:.0f} GB/s, distinct points {len(set(idx.tolist()))}, all reps {[round(x * 1e3, 1) for x in ts]} ms")
    del Q, Psi
