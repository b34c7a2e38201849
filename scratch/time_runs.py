import sys, time, subprocess, threading
sys.path.insert(0, "reducedbasismethods.jl_b200")
import numpy as np, torch
import rbm_b200 as r
n, nh, nt = 10_000_000, 64, 250
L = r.BumpOnTail.domain_length(0.3)
x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=1234)
ctx = r.Context(0)
h = r.PICHandle(n, r.PoissonSolverPBSplines(3, nh, L), ctx)
xd = torch.from_numpy(x).cuda(); vd = torch.from_numpy(v).cuda()
h.set_particles(xd, vd, float(w[0]))
Xd = torch.empty((1, 2, n), dtype=torch.float64, device="cuda")
Wd = torch.empty((1, 2), dtype=torch.float64, device="cuda")
rows = []
import os
POLL = os.environ.get("POLL", "1") == "1"
p = subprocess.Popen(["nvidia-smi" if POLL else "true", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown", "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
def rd():
    for line in p.stdout: rows.append((time.perf_counter(), line.strip()))
threading.Thread(target=rd, daemon=True).start()
time.sleep(1.0)
for rep in range(12):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h.run([1.0], 0.1, nt, 2, X=Xd, W=Wd)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    cl = [r_[1] for r_ in rows if t0 <= r_[0] <= t1]
    print(rep, f"{(t1-t0)*1e3:.2f} ms", cl[:3])
p.terminate()
