"""Back-to-back vs synchronised rbm_pic_run timing under RBM_GRAPH / RBM_PDL (diagnostic)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np, torch
import rbm_b200 as r
n, nh, nt = 10_000_000, 64, 250
L = r.BumpOnTail.domain_length(0.3)
x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=1234)
ctx = r.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
h = r.PICHandle(n, r.PoissonSolverPBSplines(3, nh, L), ctx)
h.set_particles(torch.from_numpy(x).cuda(), torch.from_numpy(v).cuda(), float(w[0]))
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 2
X = torch.empty((1, ns, n), dtype=torch.float64, device="cuda"); V = torch.empty_like(X); A = torch.empty_like(X)
W = torch.empty((1, ns), dtype=torch.float64, device="cuda"); K = torch.empty_like(W); M = torch.empty_like(W)
go = lambda: h.run([1.0], 0.1, nt, ns, X=X, V=V, A=A, W=W, K=K, M=M)
for _ in range(3): go()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter(); go(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
for _ in range(10): go()
t_host = time.perf_counter() - t0
e1.record(); torch.cuda.synchronize()
print(f"GRAPH={os.environ.get('RBM_GRAPH','1')} PDL={os.environ.get('RBM_PDL','1')} ns={ns}: synced best {min(ts)*1e3:.2f} ms; "
      f"back-to-back {e0.elapsed_time(e1)/10:.2f} ms per run (host issue time {t_host/10*1e3:.2f} ms per run)")
