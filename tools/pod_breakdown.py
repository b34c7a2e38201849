"""Where the time of one POD call goes at a tall shape: Gram (DMMA), EVD (cuSOLVER), basis formation (DMMA).
usage: python tools/pod_breakdown.py [rows] [cols] [k]"""
import os
import sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np, torch
import rbm_b200 as r
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
ctx = r.Context(0); r.set_default_context(ctx)
m = cols // 2
torch.manual_seed(0)
# low-rank-plus-noise snapshots so that the spectrum decays like a real training set
U = torch.randn(rows, 64, dtype=torch.float64, device="cuda")
def block():
    C = torch.randn(64, m, dtype=torch.float64, device="cuda") * (0.7 ** torch.arange(64, dtype=torch.float64, device="cuda"))[:, None]
    return (U @ C + 1e-6 * torch.randn(rows, m, dtype=torch.float64, device="cuda")).T.contiguous().T
X, V = block(), block()
del U
def T(f, *a, **kw):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = f(*a, **kw); torch.cuda.synchronize(); return out, time.perf_counter() - t0
for rep in range(2):
    G, tg = T(r.gram, X, V)
    (Psi, Lam), tp = T(r.get_PODBasis_cotangentLiftEVD, X, V, k=k)
    Gd = torch.from_numpy(np.ascontiguousarray(G)).cuda()
    _, te = T(torch.linalg.eigh, Gd)
    print(f"{rows} x {cols}, k={k}: gram (host result) {tg*1e3:.1f} ms | whole POD call {tp*1e3:.1f} ms | torch.linalg.eigh of G alone {te*1e3:.1f} ms")
