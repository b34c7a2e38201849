import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np, torch, torch.distributed as dist
import rbm_b200 as r
from rbm_b200 import sharding
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world, rank = dist.get_world_size(), dist.get_rank()
ctx = r.Context(local); sharding.attach_communicator(ctx)
rows_total, cols = 20_000_000, 2000                    # BASELINE configs[3]: 2e7 x 2000, row-sharded
rows = min(rows_total // world, 5_000_000)             # cap by HBM: 5e6 x 2000 x 8 B = 80 GB per rank
m = cols // 2
torch.manual_seed(rank)
X = torch.randn(m, rows, dtype=torch.float64, device="cuda").T
V = torch.randn(m, rows, dtype=torch.float64, device="cuda").T
G = torch.empty(cols, cols, dtype=torch.float64, device="cuda")
keep, P, N, Ld, nb, nrows, C = r.reducedbasis._blocks((X, V))
ts = []
for rep in range(3):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    ctx.check(ctx.lib.rbm_gram(ctx.h, P, N, Ld, nb, nrows, G.data_ptr()))
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
t = sharding.max_over_ranks(min(ts[1:]))
if rank == 0:
    R = rows * world
    print(f"distributed Gram, {world} ranks x {rows} rows x {cols} cols (global {R} x {cols}): {t*1e3:.1f} ms, "
          f"{R*cols*(cols+1)/t/1e12:.1f} TFLOP/s aggregate, {8*R*cols/t/1e9:.0f} GB/s of snapshot data")
ctx.close(); dist.destroy_process_group()
