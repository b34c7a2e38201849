import os
import sys, time, subprocess, threading
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np, torch
import rbm_b200 as r
ctx = r.Context(0)
rows_list = [1_000_000, 5_000_000]
rowsq = []
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.sw_thermal_slowdown", "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
def rd():
    for line in p.stdout: rowsq.append((time.perf_counter(), line.strip()))
threading.Thread(target=rd, daemon=True).start()
time.sleep(1.0)
for rows in rows_list:
    cols = 2000; m = cols // 2
    X = torch.randn(m, rows, dtype=torch.float64, device="cuda").T
    V = torch.randn(m, rows, dtype=torch.float64, device="cuda").T
    G = torch.empty(cols, cols, dtype=torch.float64, device="cuda")
    keep, P, N, Ld, nb, nrows, C = r.reducedbasis._blocks((X, V))
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.check(ctx.lib.rbm_gram(ctx.h, P, N, Ld, nb, nrows, G.data_ptr()))
        torch.cuda.synchronize(); t1 = time.perf_counter()
        cl = [q[1] for q in rowsq if t0 <= q[0] <= t1]
        print(f"rows={rows} rep{rep}: {(t1-t0)*1e3:.1f} ms {rows*cols*(cols+1)/(t1-t0)/1e12:.1f} TF clocks: {cl[:2]} ... {cl[-2:]}")
    del X, V, G
    torch.cuda.empty_cache()
p.terminate()
