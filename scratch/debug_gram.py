import sys, os
sys.path.insert(0, "reducedbasismethods.jl_b200")
import numpy as np, torch
import rbm_b200 as r
ctx = r.Context(0)
torch.manual_seed(0)
for n, c in [(20000, 256), (100000, 256), (300000, 256), (1000000, 256), (1000000, 130), (2000000, 64)]:
    A = torch.randn(c, n, dtype=torch.float64, device="cuda").T
    G = torch.from_numpy(r.gram(A, ctx=ctx)).cuda()
    ref = A.T @ A
    err = ((G - ref).abs().max() / ref.abs().max()).item()
    print(n, c, "gram rel err vs cuBLAS:", err, "diag ratio", (G.diag() / ref.diag()).mean().item())
n, c = 1_000_000, 256
Q, _ = torch.linalg.qr(torch.randn(n, c, dtype=torch.float64, device="cuda"))
print("QR orthogonality:", (Q.T @ Q - torch.eye(c, dtype=torch.float64, device="cuda")).abs().max().item())
