"""Debug build only (make EXTRA=-DRBM_TIMING OBJDIR=../build_tim OUT=../lib_tim/librbm_b200.so): cycles between the marks of
block 0's thread 0 along the step kernel (field solve phases, particle stream, publish).
usage: RBM_B200_LIB=.../lib_tim/librbm_b200.so python tools/time_marks.py [tiles,...] [nh]"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np, torch
import rbm_b200 as r
from rbm_b200 import _lib
lib = _lib.load()
L = r.BumpOnTail.domain_length(0.3)
ctx = r.Context(0)
nt = 1000
nh = int(sys.argv[2]) if len(sys.argv) > 2 else 64
names = {3: "wait released -> group sums", 4: "rho, mean", 5: "mat-vec", 6: "phi, quadratic", 7: "table", 8: "zero scratch",
         9: "particle stream", 10: "publish"}
for tiles in ([int(t) for t in sys.argv[1].split(",")] if len(sys.argv) > 1 else (1, 44)):
    n = 148 * 1536 * tiles
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=1)
    h = r.PICHandle(n, r.PoissonSolverPBSplines(3, nh, L), ctx)
    h.set_particles(torch.from_numpy(x).cuda(), torch.from_numpy(v).cuda(), float(w[0]))
    W = torch.empty((1, 2), dtype=torch.float64, device="cuda")
    out = (ctypes.c_double * 16)()
    for rep in range(3):
        h.run([1.0], 0.1, nt, 2, W=W)
        torch.cuda.synchronize()
        lib.rbm_debug_timing(out)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h.run([1.0], 0.1, nt, 2, W=W)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    lib.rbm_debug_timing(out)
    nl = out[0]
    print(f"tiles/block={tiles} nh={nh}: {dt/nt*1e6:.2f} us per step, {int(nl)} launches, W = {W.cpu().numpy().ravel()[-1]!r}")
    for i in range(3, 11):
        print(f"   {names[i]:30s} {out[i]/nl:9.0f} cycles")
    h.close()
