"""torchrun worker for tests/test_gpu_multi.py (one rank per GPU, NCCL).  Checks, against a single-GPU run of
the same library on rank 0 and against the CPU oracle:
  * particle-chunk sharding with the per-step NCCL all-reduce of the charge density inside librbm_b200,
  * sample partition without any collective,
  * the row-sharded Gram / POD (Gram all-reduce) and DEIM (arg-max reduction)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "reducedbasismethods.jl_b200"))

import rbm_b200 as r  # noqa: E402
from oracle import pic_oracle as po, pod_numpy as pod  # noqa: E402
from rbm_b200 import sharding  # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world, rank = dist.get_world_size(), dist.get_rank()
    L = r.BumpOnTail.domain_length(0.3)
    n, nh, nt = 200_002, 64, 12
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=21)
    particles = r.ParticleList(x[None], v[None], w[None])
    poisson = r.PoissonSolverPBSplines(3, nh, L)
    chis = np.linspace(1 / 3, 5 / 3, 5)
    IP = r.IntegratorParameters(0.1, nt, nt + 1, nh, n, len(chis))

    # ---- reference: single-GPU run without communicator (every rank does it; cheap)
    solo_ctx = r.Context(local)
    solo = r.integrate_vp_all(particles, poisson, chis, IP, ctx=solo_ctx)

    # ---- sample partition, no collective
    ctx = r.Context(local)
    ss, (slo, shi) = sharding.generate_training_set(particles, poisson, chis, IP, "samples", ctx=ctx)
    # (not bitwise: the work-item plan, hence the summation order, depends on how many samples share a launch)
    assert np.max(np.abs(ss.X - solo.X[slo:shi])) < 1e-11 and rel(ss.W, solo.W[slo:shi]) < 1e-11
    allW = sharding.gather_rows(ss.W)
    assert allW.shape == solo.W.shape and rel(allW, solo.W) < 1e-11

    # ---- particle chunks + per-step NCCL all-reduce inside the library
    cctx = r.Context(local)
    cs, (plo, phi) = sharding.generate_training_set(particles, poisson, chis, IP, "particles", ctx=cctx)
    for k in ("W", "K", "M"):
        assert rel(getattr(cs, k), getattr(solo, k)) < 1e-11, (k, rel(getattr(cs, k), getattr(solo, k)))
    assert rel(cs.Phi, solo.Phi) < 1e-10
    assert np.max(np.abs(cs.X - solo.X[:, :, plo:phi])) < 1e-11
    assert np.max(np.abs(cs.A - solo.A[:, :, plo:phi])) < 1e-9
    ref = po.integrate(x, v, float(w[0]), L, nh, float(chis[2]), 0.1, nt, nt + 1, extended=True, nthreads=4, outputs="")
    assert rel(cs.W[2], ref["W"]) < 1e-10
    # the same run with the exchange forced through ncclAllReduce (peer-memory path off) agrees to round-off
    nctx = r.Context(local)
    sharding.attach_communicator(nctx, peer_exchange=False)
    nctx._comm_attached = True
    ns_, _ = sharding.generate_training_set(particles, poisson, chis, IP, "particles", ctx=nctx)
    assert rel(ns_.W, cs.W) < 1e-12 and np.max(np.abs(ns_.X - cs.X)) < 1e-12
    nctx.close()

    # ---- chunk mode with device-resident outputs: the time loop (fused peer exchange included) is captured into a CUDA
    # graph at the second identical call and replayed from then on; all three executions agree
    hc = r.PICHandle(phi - plo, poisson, cctx)
    hc.set_particles(torch.from_numpy(x[plo:phi].copy()).cuda(), torch.from_numpy(v[plo:phi].copy()).cuda(), float(w[0]))
    runs = []
    for rep in range(4):
        Wd = torch.zeros((len(chis), nt + 1), dtype=torch.float64, device="cuda")
        Kd = torch.zeros_like(Wd); Md = torch.zeros_like(Wd)
        l0 = cctx.launch_count()
        hc.run(chis, 0.1, nt, nt + 1, W=Wd, K=Kd, M=Md)
        runs.append((Wd.cpu().numpy(), Kd.cpu().numpy(), Md.cpu().numpy(), cctx.launch_count() - l0))
    for rep in (1, 2, 3):
        assert np.array_equal(runs[rep][0], runs[0][0]), "graph replay of the chunk-mode loop changed W"
        assert rel(runs[rep][1], runs[0][1]) < 1e-13 and rel(runs[rep][2], runs[0][2]) < 1e-12
    assert rel(runs[0][0], cs.W) < 1e-12 and rel(runs[0][1], cs.K) < 1e-12
    os.environ["RBM_PEER_XCH"] = "0"                       # the NCCL exchange inside a captured loop
    for rep in range(3):
        Wn = torch.zeros((len(chis), nt + 1), dtype=torch.float64, device="cuda")
        hc.run(chis, 0.1, nt, nt + 1, W=Wn)
        assert rel(Wn.cpu().numpy(), runs[0][0]) < 1e-12
    del os.environ["RBM_PEER_XCH"]
    Wp = torch.zeros((len(chis), nt + 1), dtype=torch.float64, device="cuda")
    hc.run(chis, 0.1, nt, nt + 1, W=Wp)                    # back on the peer path (new graph key)
    assert np.array_equal(Wp.cpu().numpy(), runs[0][0])
    hc.close()

    # ---- row-sharded POD: Gram all-reduce, replicated EVD, local basis rows, DEIM arg-max reduction
    Xl, Vl, Al = cs.matrix("X"), cs.matrix("V"), cs.matrix("A")
    Xg, Vg, Ag = solo.matrix("X"), solo.matrix("V"), solo.matrix("A")
    G = r.gram(Xl, Vl, ctx=cctx)
    assert rel(G, pod.gram(Xg, Vg)) < 1e-11
    assert np.array_equal(G, G.T), "the all-reduced Gram must be exactly symmetric (it is mirrored after the reduction)"
    Psi_l, Lam = r.get_PODBasis_EVD(Al, k=12, ctx=cctx)
    Psi_g, Lam_g = r.get_PODBasis_EVD(Ag, k=12, ctx=solo_ctx)
    lead = Lam_g / Lam_g[0] > 1e-6
    assert np.max(np.abs(Lam[lead] - Lam_g[lead]) / Lam_g[lead]) < 1e-10
    full = sharding.gather_rows(Psi_l)
    for j in range(6):
        assert min(np.linalg.norm(full[:, j] - Psi_g[:, j]), np.linalg.norm(full[:, j] + Psi_g[:, j])) < 1e-6
    idx = r.deim_indices(Psi_l, ctx=cctx)                       # global row numbers
    assert np.array_equal(idx, pod.deim_indices(full))
    Z = r.project(Psi_l, Al, ctx=cctx)
    assert rel(Z, full.T @ Ag) < 1e-10
    # ---- reduced-model sweep: samples partitioned over ranks, traces gathered in sample order
    n_r = 3000
    xr, vr, wr = r.BumpOnTail.draw_accept_reject(n_r, seed=9)
    pr = r.ParticleList(xr[None], vr[None], wr[None]); poi = r.PoissonSolverPBSplines(3, 16, L)
    train = np.array([0.7, 1.0, 1.3])
    ipt = r.IntegratorParameters(0.1, 20, 21, 16, n_r, 3)
    tsr = r.TrainingSet.create(pr, poi, 21, {"kappa": 0.3}, r.ParameterSpace(r.Parameter("chi", train)), ipt)
    r.integrate_vp_all(pr, poi, train, ipt, tsr.snapshots, ctx=solo_ctx)
    rb = r.ReducedBasis.from_trainingset(r.CotangentLiftEVD(), tsr, particle_tol=1e-8, field_tol=1e-6, ctx=solo_ctx)
    rfield = r.DEIMElectricField(poi, rb.Psi_p, rb.Psi_e, rb.Pi_e)
    sweep = np.linspace(0.75, 1.25, 5)
    ips = r.IntegratorParameters(0.1, 20, 21, 16, n_r, len(sweep))
    rss_l, (slo, shi), traces = sharding.reduced_sweep(pr, rfield, sweep, ips, ctx=solo_ctx)
    rss_all = r.reduced_integrate_vp_all(pr, rfield, sweep, ips, ctx=solo_ctx)
    assert traces["W"].shape == (len(sweep), 21)
    assert rel(traces["W"], rss_all.W) < 1e-11 and rel(rss_l.X, rss_all.X[slo:shi]) < 1e-11
    # ---- row-sharded reduced model: basis rows and particles split like the particle chunks, rho summed over ranks per
    # step, the small reduced operators formed by all-reduce at create time; equals the single-rank model
    rlo, rhi = sharding.partition_particles(n_r, world, rank)
    prl = r.ParticleList(pr.x[:, rlo:rhi], pr.v[:, rlo:rhi], pr.w[:, rlo:rhi])
    idx_global = np.argmax(rb.host("Pi_e"), axis=0)
    rfield_l = r.DEIMElectricField(poi, np.asfortranarray(rb.host("Psi_p")[rlo:rhi]), np.asfortranarray(rb.host("Psi_e")[rlo:rhi]),
                                   idx_global)
    ipl = r.IntegratorParameters(0.1, 20, 21, 16, rhi - rlo, len(sweep))
    rss_sh = r.reduced_integrate_vp_all(prl, rfield_l, sweep, ipl, ctx=cctx)
    for k in ("W", "K", "M"):
        assert rel(getattr(rss_sh, k), getattr(rss_all, k)) < 1e-9, (k, rel(getattr(rss_sh, k), getattr(rss_all, k)))
    assert rel(rss_sh.Phi, rss_all.Phi) < 1e-9
    assert np.max(np.abs(rss_sh.X - rss_all.X[:, :, rlo:rhi])) < 1e-9
    if rank == 0:
        print(f"multi-GPU checks ok on {world} GPUs")
    for c in (cctx, ctx, solo_ctx):
        c.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
