import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np, torch, torch.distributed as dist
import rbm_b200 as r
from rbm_b200 import sharding
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world, rank = dist.get_world_size(), dist.get_rank()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000   # particles PER RANK (weak scaling of the particle set)
nsamp = int(sys.argv[2]) if len(sys.argv) > 2 else 1          # parameter samples advanced in lock-step on every rank
nh, nt = 64, 250
L = r.BumpOnTail.domain_length(0.3)
x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=1234 + rank)
ctx = r.Context(local)
sharding.attach_communicator(ctx, peer_exchange=False)
peer = sharding.enable_peer_exchange(ctx) if os.environ.get('RBM_PEER_XCH', '1') == '1' else False
if rank == 0: print('peer exchange enabled:', peer)
h = r.PICHandle(n, r.PoissonSolverPBSplines(3, nh, L), ctx)
h.set_particles(torch.from_numpy(x).cuda(), torch.from_numpy(v).cuda(), L / (n * world))
Wd = torch.empty((nsamp, 2), dtype=torch.float64, device="cuda")
chis = np.linspace(1 / 3, 5 / 3, nsamp) if nsamp > 1 else np.array([1.0])
ts = []
for rep in range(8):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    h.run(chis, 0.1, nt, 2, W=Wd)
    torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
best = sharding.max_over_ranks(min(ts[2:]))
if rank == 0:
    print(f"chunk mode, {world} ranks x {n} particles, {nsamp} sample(s): {best:.2f} ms per {nt} steps = {best/nt*1e3:.1f} us/step, {world*n*nsamp*nt/best*1e3:.3e} particle-steps/s, W={Wd.cpu().numpy()[0,-1]:.6f}")
h.close(); ctx.close(); dist.destroy_process_group()
