"""Multi-GPU host logic: one process per GPU (torchrun), torch.distributed for the plumbing.

Two ways to split the training-set generation of scripts/bump_on_tail_1_training.jl:87-109 over ranks
(SURVEY.md 8(e)):

* ``samples``   -- every rank integrates its own contiguous block of parameter samples with ALL particles;
                   samples are independent (the reference restarts each one from the same `particles`), so
                   there is no data-path collective.  This is what bench.py scales (weak scaling).
* ``particles`` -- every rank holds a contiguous chunk of the particles and integrates ALL samples; the
                   charge density (nparam x nh doubles) and the diagnostics K, M are summed over ranks once
                   per step by NCCL inside librbm_b200 (rbm_comm_init).  Snapshots come out row-sharded, which
                   is what the POD wants: the Gram matrix is additive over row blocks (one all-reduce).

Everything here that does not touch the GPU (partitions, id broadcast, gathers) works on any
torch.distributed backend and is covered by the world_size-2 gloo tests.
"""
from __future__ import annotations

import numpy as np


def partition(n: int, world: int, rank: int):
    """Contiguous balanced block [lo, hi) of n items for `rank` (first n % world ranks get one more)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(int(n), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def partition_particles(n: int, world: int, rank: int, align: int = 4):
    """Like partition() but chunk starts are multiples of `align` particles (16/32-byte aligned vector loads)."""
    blocks = (int(n) + align - 1) // align
    lo, hi = partition(blocks, world, rank)
    return min(lo * align, n), min(hi * align, n)


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0, group=None) -> bytes:
    """Broadcast a fixed-size byte string (the 128-byte NCCL unique id) from `src` to every rank."""
    import torch
    import torch.distributed as dist

    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if dist.get_rank(group) == src:
        if payload is None or len(payload) != nbytes:
            raise ValueError("source rank must supply the payload")
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(buf, src=src, group=group)
    return bytes(buf.cpu().numpy().tobytes())


def attach_communicator(ctx, group=None, peer_exchange: bool = True):
    """Give an rbm_ctx an NCCL communicator spanning the torch.distributed group (rank 0 creates the id)."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return
    uid = ctx.comm_unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, 128, 0, group)
    ctx.comm_init(world, rank, uid)
    if peer_exchange:
        enable_peer_exchange(ctx, group)


def all_gather_bytes(payload: bytes, group=None) -> bytes:
    """Concatenation, in rank order, of one fixed-size byte string per rank."""
    import torch
    import torch.distributed as dist

    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    parts = [torch.empty_like(mine) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, mine, group=group)
    return b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)


def enable_peer_exchange(ctx, group=None) -> bool:
    """Switch the per-step charge-density exchange from ncclAllReduce to direct NVLink peer loads: every rank
    exposes its exchange buffer through CUDA IPC, the 64-byte handles are all-gathered, and the path is used only
    if EVERY rank could map all its peers (otherwise all ranks stay on NCCL)."""
    import torch
    import torch.distributed as dist

    handles = all_gather_bytes(ctx.comm_peer_handle(), group)
    ok = 1
    try:
        ctx.comm_peer_attach(handles)
    except Exception:
        ok = 0
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    flag = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:
        ctx.comm_peer_attach(None)
        return False
    return True


def gather_rows(local: np.ndarray, group=None) -> np.ndarray:
    """Concatenate per-rank arrays along axis 0 on every rank (ragged first dimension allowed)."""
    import torch.distributed as dist

    objs = [None] * dist.get_world_size(group)
    dist.all_gather_object(objs, np.ascontiguousarray(local), group=group)
    return np.concatenate(objs, axis=0)


def max_over_ranks(value: float, group=None) -> float:
    """The slowest rank's time: how every multi-GPU number in this repo is reduced."""
    import torch
    import torch.distributed as dist

    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def generate_training_set(particles, poisson, chis, IP, mode: str = "samples", ctx=None, group=None):
    """Distributed counterpart of integrator.integrate_vp_all.

    mode == "samples":   returns this rank's Snapshots for samples [lo, hi) and the slice (lo, hi).
    mode == "particles": `particles` is the GLOBAL ParticleList; this rank integrates its chunk of it for all
                         samples; returns Snapshots whose X, V, A hold the chunk's rows (Phi, W, K, M are global
                         and identical on every rank) and the particle slice (lo, hi).
    """
    import torch.distributed as dist

    from .engine import default_context
    from .integrator import integrate_vp_all
    from .particles import ParticleList
    from .snapshots import IntegratorParameters

    ctx = ctx or default_context()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    chis = np.ascontiguousarray(chis, dtype=np.float64).reshape(-1)
    if mode == "samples":
        lo, hi = partition(chis.shape[0], world, rank)
        ip = IntegratorParameters(IP.dt, IP.nt, IP.ns, IP.nh, IP.np, hi - lo)
        return integrate_vp_all(particles, poisson, chis[lo:hi], ip, ctx=ctx), (lo, hi)
    if mode == "particles":
        if world > 1 and not getattr(ctx, "_comm_attached", False):
            attach_communicator(ctx, group)
            ctx._comm_attached = True
        lo, hi = partition_particles(len(particles), world, rank)
        local = ParticleList(particles.x[:, lo:hi], particles.v[:, lo:hi], particles.w[:, lo:hi])
        ip = IntegratorParameters(IP.dt, IP.nt, IP.ns, IP.nh, hi - lo, chis.shape[0])
        return integrate_vp_all(local, poisson, chis, ip, ctx=ctx), (lo, hi)
    raise ValueError("mode must be 'samples' or 'particles'")


def reduced_sweep(particles, rfield, chis, IP, ctx=None, group=None, gather_traces: bool = True):
    """Distributed counterpart of reduced.reduced_integrate_vp_all for a parameter sweep (BASELINE configs[4]): the
    test samples are independent (scripts/bump_on_tail_3_testing.jl:108-138 loops over them), so each rank advances its
    slice [lo, hi) of `chis` with the full basis and no data-path collective.  Returns (local Snapshots, (lo, hi),
    traces) where traces = dict(W, K, M) gathered over all samples in order on every rank (None if not asked)."""
    import torch.distributed as dist

    from .engine import default_context
    from .reduced import reduced_integrate_vp_all
    from .snapshots import IntegratorParameters

    ctx = ctx or default_context()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    chis = np.ascontiguousarray(chis, dtype=np.float64).reshape(-1)
    lo, hi = partition(chis.shape[0], world, rank)
    ip = IntegratorParameters(IP.dt, IP.nt, IP.ns, IP.nh, IP.np, hi - lo)
    rss = reduced_integrate_vp_all(particles, rfield, chis[lo:hi], ip, ctx=ctx)
    traces = None
    if gather_traces:
        if world > 1:
            traces = {k: gather_rows(getattr(rss, k), group) for k in ("W", "K", "M")}
        else:
            traces = {k: getattr(rss, k) for k in ("W", "K", "M")}
    return rss, (lo, hi), traces
