"""Debug build only (-DRBM_TIMING): per-CTA timeline of the last step-kernel launch (np = 1e7, nh = 64): which SMs finish
their chunk late.  usage: RBM_B200_LIB=.../lib_tim/librbm_b200.so python tools/time_blocks.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np, torch
import rbm_b200 as r
from rbm_b200 import _lib
lib = _lib.load()
L = r.BumpOnTail.domain_length(0.3)
ctx = r.Context(0)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=1)
h = r.PICHandle(n, r.PoissonSolverPBSplines(3, 64, L), ctx)
h.set_particles(torch.from_numpy(x).cuda(), torch.from_numpy(v).cuda(), float(w[0]))
W = torch.empty((1, 2), dtype=torch.float64, device="cuda")
for nt in (100, 101, 150, 151):
    h.run([1.0], 0.1, nt, 2, W=W); torch.cuda.synchronize()
    h.run([1.0], 0.1, nt, 2, W=W); torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (148 * 6))()
    lib.rbm_debug_blocks(buf, 148)
    a = np.array(buf[:], dtype=np.int64).reshape(148, 6)
    smid = a[:, 0]; t0 = a[:, 1].min()
    rel = (a[:, 1:5] - t0) / 1e3
    dur = rel[:, 2] - rel[:, 1]
    order = np.argsort(dur)
    print(f"nt={nt}: wait released {rel[:,0].min():.2f}..{rel[:,0].max():.2f} us, stream start {rel[:,1].min():.2f}..{rel[:,1].max():.2f}, "
          f"stream end {rel[:,2].min():.2f}..{rel[:,2].max():.2f}, publish end {rel[:,3].min():.2f}..{rel[:,3].max():.2f}")
    print(f"   stream duration us: min {dur.min():.2f} p25 {np.percentile(dur,25):.2f} median {np.median(dur):.2f} p75 {np.percentile(dur,75):.2f} max {dur.max():.2f}")
    print("   slowest 12 (block:smid:us):", " ".join(f"{b}:{smid[b]}:{dur[b]:.1f}" for b in order[-12:]))
    print("   fastest 12 (block:smid:us):", " ".join(f"{b}:{smid[b]}:{dur[b]:.1f}" for b in order[:12]))
    if nt == 151:
        print("   by smid:", " ".join(f"{smid[b]}:{dur[b]:.1f}" for b in np.argsort(smid)))
