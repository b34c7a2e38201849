"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path (partitions, NCCL-id broadcast plumbing,
gathers, max-over-ranks timing).  The data-path collective itself (NCCL inside librbm_b200) is exercised on
real GPUs by tests/test_gpu_multi.py.  The sharded-deposit identity the library relies on (charge density is
additive over particle chunks; samples are independent) is checked here with the CPU oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rbm_b200 import sharding


def test_partition_properties():
    for n in (0, 1, 7, 64, 10_000_001):
        for world in (1, 2, 3, 8):
            blocks = [sharding.partition(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
            pb = [sharding.partition_particles(n, world, r) for r in range(world)]
            assert pb[0][0] == 0 and pb[-1][1] == n and all(lo % 4 == 0 or lo == n for lo, _ in pb)
            assert all(pb[i][1] == pb[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        sharding.partition(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "reducedbasismethods.jl_b200"))
    from oracle import pic_oracle as po
    from rbm_b200.bump_on_tail import domain_length, draw_accept_reject

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. the 128-byte id travels from rank 0 to everyone, byte for byte
        uid = bytes(range(128)) if rank == 0 else None
        got = sharding.broadcast_bytes(uid, 128, 0)
        assert got == bytes(range(128))
        # 2. particle-chunk sharding: per-rank deposits summed over ranks == the global deposit
        L = domain_length(0.3); nh = 16; n = 10_001
        x, v, w = draw_accept_reject(n, seed=3)
        lo, hi = sharding.partition_particles(n, world, rank)
        local = torch.from_numpy(po.deposit(x[lo:hi], float(w[0]), L, nh))
        dist.all_reduce(local)                                   # what ncclAllReduce does per step on the GPUs
        full = po.deposit(x, float(w[0]), L, nh)
        assert np.max(np.abs(local.numpy() - full)) < 1e-13 * np.max(np.abs(full))
        # 3. sample partition: each rank integrates its own samples; gathering restores the sample order
        chis = np.linspace(1 / 3, 5 / 3, 5)
        slo, shi = sharding.partition(len(chis), world, rank)
        mine = np.stack([po.integrate(x[:500], v[:500], float(w[0]), L, nh, float(c), 0.1, 4, 5)["W"] for c in chis[slo:shi]])
        allW = sharding.gather_rows(mine)
        ref = np.stack([po.integrate(x[:500], v[:500], float(w[0]), L, nh, float(c), 0.1, 4, 5)["W"] for c in chis])
        assert allW.shape == (5, 5) and np.array_equal(allW, ref)
        # 3b. fixed-size byte strings (the 64-byte IPC handles) all-gathered in rank order
        blob = sharding.all_gather_bytes(bytes([rank]) * 64)
        assert blob == bytes([0]) * 64 + bytes([1]) * 64
        # 4. timing reduction: the slowest rank counts
        assert sharding.max_over_ranks(1.0 + rank) == float(world)
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
