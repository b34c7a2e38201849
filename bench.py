#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native hot path (contract: see the task brief / DESIGN.md section 6).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference path (oracle port)

Workload = BASELINE.json configs[1]: bump-on-tail 1D1V PIC, np = 1e7 particles, cubic B-splines with nh = 64
modes, ONE parameter sample per GPU, dt = 0.1, nt = 250 steps, ns = 2 saved steps (initial + final).
One bench "step" = one complete rbm_pic_run of that workload (nt PIC time steps).  Metric = particle-steps/s
(FP64) = n_gpus * nt * np / time.  With N > 1 (torchrun, one rank per GPU) `value` partitions the samples across
ranks (independent objects, no data-path collective): weak scaling.

  value      inputs and outputs resident in HBM (device pointers through the C ABI), CUDA events, max over ranks
  e2e        the same call with HOST (pinned) buffers: H2D of x, v and D2H of X, V, A, W, K, M inside the timed region
             (N = 1: also with pageable NumPy arrays -- what a Julia Array is -- and with the same arrays page-locked
             in place by rbm_host_register)
  roofline   k_step: 32 B per particle-step / its mean launch duration IN THE PRODUCTION RUN (graph replay + programmatic
             dependent launch): (timed bench step - the other kernels of the step) / launches; the per-launch figure of
             the serialised profile pass is kept as a second field
  save_every_step   N = 1: the shipped script's cadence ns = nt + 1 (scripts/bump_on_tail_1_training.jl:26-48), 56 B per
             particle-step, second roofline object
  reduced_sweep     N = 1: BASELINE configs[4], 256-sample reduced model with DEIM (scripts/bump_on_tail_3_testing.jl:45-57 shape)
  multi      N > 1: the COMMUNICATING paths under the same clock -- BASELINE configs[2] in particle-chunk mode (64 samples x
             1e7 particles, charge density summed over ranks every step: fused NVLink peer exchange inside k_step, and
             ncclAllReduce with RBM_PEER_XCH=0), its W/K/M traces checked against a single-rank run of the same
             particles, and BASELINE configs[3], the row-sharded Gram with one NCCL all-reduce
  cpu_baseline / --impl reference   the oracle port (oracle/pic_oracle.c, OpenMP) on the host cores, bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "reducedbasismethods.jl_b200"))

import numpy as np  # noqa: E402

KAPPA = 0.3
METRIC = "particle-steps/sec (FP64)"
UNIT = "particle-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--np", type=int, default=10_000_000, dest="n")
    ap.add_argument("--nh", type=int, default=64)
    ap.add_argument("--nt", type=int, default=250)
    ap.add_argument("--ns", type=int, default=2)
    ap.add_argument("--no-pod", action="store_true", help="skip the POD Gram side measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip save_every_step / host-buffer variants / reduced_sweep (N = 1)")
    ap.add_argument("--no-multi", action="store_true", help="skip the communicating multi-GPU section (N > 1)")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--multi-samples", type=int, default=64)
    ap.add_argument("--gram-rows", type=int, default=20_000_000)
    return ap.parse_args()


def workload_name(a):
    return (f"BASELINE configs[1]: bump-on-tail 1D1V PIC, np={a.n:.0e} particles, cubic B-splines nh={a.nh}, "
            f"1 parameter sample per GPU, dt=0.1, nt={a.nt}, ns={a.ns}")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.  The process is started long
    before the timed region (nvidia-smi start-up itself perturbs the GPU for tens of ms) and its rows are
    time-stamped on arrival; summary() keeps only the rows that fall inside [t0, t1]."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=200):
        # 200 ms is the profiling recipe's period; a 25 ms poll was measured to cost the benchmarked process ~25 % on some
        # hosts (every nvidia-smi query briefly holds the driver)
        self.index = index
        self.period_ms = period_ms
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period_ms), "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass

    def summary(self, t0, t1):
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 8]
        def num(s):
            try:
                return float(s)
            except ValueError:
                return None
        sm = [num(r[1]) for r in rows if num(r[1]) is not None]
        mx = [num(r[2]) for r in rows if num(r[2]) is not None]
        pw = [num(r[3]) for r in rows if num(r[3]) is not None]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which must not cap the CPU arm)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_oracle_rate(x, v, w, L, nh, seconds, nthreads):
    """Oracle port (literal double accumulation = the reference's arithmetic) on a bounded sample:
    the same particles and grid, fewer time steps."""
    from oracle import pic_oracle as po

    t0 = time.perf_counter()
    po.integrate(x, v, w, L, nh, 1.0, 0.1, 1, 2, extended=False, nthreads=nthreads, outputs="")
    t1 = time.perf_counter() - t0           # 1 step + 2 save-deposits (~ 2 steps of work; also warms the thread pool)
    nt = int(max(2, min(250, seconds / max(t1 / 2.0, 1e-3))))
    t0 = time.perf_counter()
    po.integrate(x, v, w, L, nh, 1.0, 0.1, nt, 2, extended=False, nthreads=nthreads, outputs="")
    dt = time.perf_counter() - t0
    return nt * x.shape[0] / dt, nt, dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import rbm_b200.bump_on_tail as bot
    from oracle import pic_oracle as po

    L = bot.domain_length(KAPPA)
    x, v, w = bot.draw_accept_reject(a.n, seed=1234)
    nthreads = host_threads()
    nt_s = max(2, min(a.nt, 20))    # bounded sample: 20 of nt steps (+ the 2 save deposits) per bench step
    for _ in range(a.warmup):
        po.integrate(x, v, float(w[0]), L, a.nh, 1.0, 0.1, 1, 2, extended=False, nthreads=nthreads, outputs="")
    t0 = time.perf_counter()
    for _ in range(a.steps):
        po.integrate(x, v, float(w[0]), L, a.nh, 1.0, 0.1, nt_s, 2, extended=False, nthreads=nthreads, outputs="")
    el = time.perf_counter() - t0
    value = a.steps * nt_s * a.n / el
    sample = f"same particles and grid, {nt_s} of {a.nt} time steps per bench step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * el / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "note": "CPU restatement of the reference path (oracle port, OpenMP); "
                   "the Julia reference itself cannot run here (no Julia; PIC arithmetic in un-vendored packages)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cublas_dgemm_tflops(torch):
    """FP64 denominator (MEASURED_PEAKS.json has only bf16): cuBLAS DGEMM 8192^3, best of 5, in this process."""
    Aq = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(Aq, Aq); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    del Aq
    torch.cuda.empty_cache()
    return 2 * 8192 ** 3 / best / 1e12


def pod_side_measurement(ctx, torch, dgemm_tf, rows=1_000_000, cols=2000):
    """POD snapshot Gram (BASELINE metric, second half): GB/s of snapshot data and FP64 TFLOP/s of the DMMA kernel
    on a tall matrix [X V] (two column blocks), against a cuBLAS DGEMM measured in the same process."""
    import rbm_b200 as r

    free = torch.cuda.mem_get_info()[0]
    while rows * cols * 8 * 1.3 > free and rows > 100_000:
        rows //= 2
    m = cols // 2
    X = torch.randn(m, rows, dtype=torch.float64, device="cuda").T      # column-major (rows, m)
    V = torch.randn(m, rows, dtype=torch.float64, device="cuda").T
    G = torch.empty(cols, cols, dtype=torch.float64, device="cuda")
    keep, P, N, Ld, nb, nrows, C = r.reducedbasis._blocks((X, V))

    def go():
        ctx.check(ctx.lib.rbm_gram(ctx.h, P, N, Ld, nb, nrows, G.data_ptr()))
    go()
    ctx.profile_enable(True)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); go(); go(); t1.record(); torch.cuda.synchronize()
    wall = t0.elapsed_time(t1) / 2 * 1e-3
    n, ms = ctx.profile_query("k_dgemm")
    ctx.profile_enable(False)
    kern = ms / max(n, 1) * 1e-3
    flops = rows * cols * (cols + 1)
    del X, V, G
    torch.cuda.empty_cache()
    return {"workload": f"Gram of [X V], {rows} rows x {cols} columns (two column blocks), FP64 DMMA",
            "snapshot_gbs": 8 * rows * cols / wall / 1e9, "tflops": flops / wall / 1e12,
            "kernel_tflops": flops / kern / 1e12, "ms": wall * 1e3, "kernel_ms": kern * 1e3,
            "cublas_dgemm_tflops_measured": dgemm_tf, "frac_of_measured_dgemm": flops / kern / 1e12 / dgemm_tf,
            "dmma_peak_tflops_microbench": 37.1,
            "roofline": {"bound": "tensor", "unit": "TFLOP/s", "achieved": flops / kern / 1e12, "peak": dgemm_tf,
                         "frac": flops / kern / 1e12 / dgemm_tf,
                         "note": "FP64 has no entry in MEASURED_PEAKS.json: peak = cuBLAS DGEMM measured in this process"}}


def timed_runs(torch, fn, reps, sync_world=None):
    """mean ms of `reps` back-to-back calls (CUDA events on the current stream; max over ranks when asked)."""
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if sync_world is not None:
        sync_world.barrier()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if sync_world is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        sync_world.all_reduce(t, op=sync_world.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def save_every_step_measurement(a, h, chi, torch, peak):
    """Second run of SURVEY 8(d) cfg2: ns = nt + 1, the cadence of the shipped script.  Device-resident snapshots
    (3 x np x ns doubles = 60 GB at np = 1e7).  Every step is one k_step<DUAL> launch: push + snapshot stores + the
    deposit of the next half-drifted positions + the deposit at the saved positions (second histogram)."""
    ns = a.nt + 1
    need = 3 * 8 * a.n * ns
    free = torch.cuda.mem_get_info()[0]
    if need > 0.8 * free:
        return {"skipped": f"needs {need / 1e9:.0f} GB of device snapshots, {free / 1e9:.0f} GB free"}
    X = torch.empty((1, ns, a.n), dtype=torch.float64, device="cuda"); V = torch.empty_like(X); A = torch.empty_like(X)
    W = torch.empty((1, ns), dtype=torch.float64, device="cuda"); K = torch.empty_like(W); M = torch.empty_like(W)

    def go():
        h.run(chi, 0.1, a.nt, ns, X=X, V=V, A=A, W=W, K=K, M=M)
    for _ in range(3):
        go()                                   # direct, capture, replay
    ms = timed_runs(torch, go, 3)
    byt = a.n * (32.0 * a.nt + 24.0 * ns)
    out = {"workload": f"BASELINE configs[1] with ns = nt + 1 = {ns} (every step saved, device-resident snapshots)",
           "value": a.nt * a.n / (ms * 1e-3), "unit": UNIT, "ms_per_run": ms, "us_per_pic_step": 1e3 * ms / a.nt,
           "roofline": {"bound": "hbm", "kernel": "k_step<SAVE=1,DUAL=1>", "unit": "GB/s",
                        "achieved": byt / (ms * 1e-3) / 1e9, "peak": peak, "frac": byt / (ms * 1e-3) / 1e9 / peak,
                        "algorithmic_bytes_per_launch": 56.0 * a.n,
                        "note": "32 B (x, v read + written) + 24 B (X, V, A stored) per particle-step; whole run / timed run"},
           "W_final": float(W[0, -1].item())}
    del X, V, A
    torch.cuda.empty_cache()
    return out


def host_buffer_variants(a, h, ctx, chi, x, v, wu, torch, W_ref):
    """e2e through pageable NumPy arrays (what a Julia Array is) and through the same arrays page-locked in place with
    rbm_host_register (what the Julia glue's `pin = true` does)."""
    Xp = np.empty((1, a.ns, a.n)); Vp = np.empty_like(Xp); Ap = np.empty_like(Xp)
    Wp = np.empty((1, a.ns)); Kp = np.empty_like(Wp); Mp = np.empty_like(Wp)
    xs, vs = x.copy(), v.copy()

    def go():
        h.set_particles(xs, vs, wu)
        h.run(chi, 0.1, a.nt, a.ns, X=Xp, V=Vp, A=Ap, W=Wp, K=Kp, M=Mp)
    out = {}
    go()
    out["pageable"] = {"value": a.nt * a.n / (timed_runs(torch, go, 2) * 1e-3), "unit": UNIT}
    assert abs(float(Wp[0, -1]) - W_ref) <= 1e-9 * abs(W_ref)
    bufs = (Xp, Vp, Ap, xs, vs)
    t0 = time.perf_counter()
    for b in bufs:
        ctx.host_register(b)
    t_reg = time.perf_counter() - t0
    try:
        go()
        out["registered"] = {"value": a.nt * a.n / (timed_runs(torch, go, 2) * 1e-3), "unit": UNIT,
                             "register_ms_once": 1e3 * t_reg}
        assert abs(float(Wp[0, -1]) - W_ref) <= 1e-9 * abs(W_ref)
    finally:
        for b in bufs:
            ctx.host_unregister(b)
    del Xp, Vp, Ap
    # training-set generation as the reference's script does it (bump_on_tail_1_training.jl:87-109: every sample starts from
    # the same particles and is copied into the snapshot arrays): ONE call for several samples with pinned host arrays --
    # one H2D of x, v, and every group's snapshot copies run under the next group's time loop (two staging sets)
    S = 4
    xh = torch.from_numpy(x).pin_memory(); vh = torch.from_numpy(v).pin_memory()
    Xs = torch.empty((S, a.ns, a.n), dtype=torch.float64).pin_memory(); Vs = torch.empty_like(Xs).pin_memory()
    As = torch.empty_like(Xs).pin_memory(); Ws = torch.empty((S, a.ns), dtype=torch.float64).pin_memory()
    chis = np.linspace(1.0 / 3.0, 5.0 / 3.0, S)

    def go_set():
        h.set_particles(xh, vh, wu)
        h.run(chis, 0.1, a.nt, a.ns, X=Xs, V=Vs, A=As, W=Ws)
    go_set()
    out["trainingset"] = {"value": S * a.nt * a.n / (timed_runs(torch, go_set, 2) * 1e-3), "unit": UNIT, "samples_per_call": S,
                          "h2d_bytes_per_call": 16 * a.n, "d2h_bytes_per_call": S * 24 * a.n * a.ns,
                          "host_buffers": "pinned (torch)"}
    return out


def reduced_sweep_measurement(ctx, torch, dgemm_tf, n=100_000, nh=16, nt=250, nsamp=256):
    """BASELINE configs[4]: reduced model with DEIM over a 256-sample sweep (scripts/bump_on_tail_3_testing.jl:45-57 with 256
    samples), basis from a shipped-default-shaped training set (10 samples, every step saved) kept in HBM."""
    import rbm_b200 as r

    L = r.BumpOnTail.domain_length(KAPPA)
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=1234)
    particles = r.ParticleList(x[None], v[None], w[None]); poisson = r.PoissonSolverPBSplines(3, nh, L)
    train = np.linspace(1 / 3, 5 / 3, 10)
    ip = r.IntegratorParameters(0.1, nt, nt + 1, nh, n, 10)
    ts = r.TrainingSet.create(particles, poisson, nt + 1, {"kappa": KAPPA}, r.ParameterSpace(r.Parameter("chi", train)), ip,
                              device="cuda")
    t0 = time.perf_counter()
    r.integrate_vp_all(particles, poisson, train, ip, ts.snapshots, ctx=ctx)
    rb = r.ReducedBasis.from_trainingset(r.CotangentLiftEVD(), ts, particle_tol=1e-9, field_tol=1e-5, ctx=ctx)
    torch.cuda.synchronize()
    t_offline = time.perf_counter() - t0
    rfield = r.DEIMElectricField(poisson, rb.Psi_p, rb.Psi_e, rb.Pi_e)
    chis = np.sort(np.random.default_rng(1234).uniform(1 / 3, 5 / 3, nsamp))
    model = r.ReducedModel(particles, rfield, ctx)
    W = torch.zeros((nsamp, 2), dtype=torch.float64, device="cuda")

    def go():
        model.run(chis, 0.1, nt, 2, W=W)
    go()
    ms = timed_runs(torch, go, 2)
    flops = 2.0 * n * rb.k_p * nsamp * nt                 # X = Psi_p Z for all samples, every step (the dominant product)
    out = {"workload": f"BASELINE configs[4]: reduced model + DEIM, {nsamp}-sample sweep, np={n}, nh={nh}, nt={nt}, "
                       f"k_p={rb.k_p}, k_e={rb.k_e} (basis from 10 training samples x {nt + 1} snapshots)",
           "ms_per_sweep": ms, "us_per_step": 1e3 * ms / nt, "sample_steps_per_s": nsamp * nt / (ms * 1e-3),
           "particle_steps_per_s": nsamp * nt * n / (ms * 1e-3), "offline_stage_s": t_offline,
           "roofline": {"bound": "tensor", "unit": "TFLOP/s", "kernel": "k_recon_deposit (X = Psi Z with the deposit as its epilogue)",
                        "achieved": flops / (ms * 1e-3) / 1e12, "peak": dgemm_tf, "frac": flops / (ms * 1e-3) / 1e12 / dgemm_tf,
                        "note": "2 N k_p S flop per step over the WHOLE timed step (deposit, solve, DEIM gather included); "
                                "peak = cuBLAS DGEMM measured in this process"}}
    model.close()
    del ts, rb, rfield, W
    torch.cuda.empty_cache()
    return out


def multi_gpu_section(a, torch, dist, r, x, v, wu, L, value_sample_partition, dgemm_tf):
    """The communicating multi-GPU paths, measured in the same driver run as `value` (which has no collective)."""
    from rbm_b200 import sharding

    world, rank = dist.get_world_size(), dist.get_rank()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    out = {}
    nsamp, nt, ns = a.multi_samples, a.nt, 6
    chis = np.linspace(1 / 3, 5 / 3, nsamp)
    lo, hi = sharding.partition_particles(a.n, world, rank)
    poisson = r.PoissonSolverPBSplines(3, a.nh, L)

    # ---- (a) BASELINE configs[2]: 64 samples x 1e7 particles, particle chunks over the ranks
    cctx = r.Context(local)
    cctx.set_stream(torch.cuda.current_stream().cuda_stream)
    sharding.attach_communicator(cctx, peer_exchange=False)
    peer = sharding.enable_peer_exchange(cctx)
    hc = r.PICHandle(hi - lo, poisson, cctx)
    hc.set_particles(torch.from_numpy(x[lo:hi].copy()).cuda(), torch.from_numpy(v[lo:hi].copy()).cuda(), wu)
    # timed with the save cadence of the headline workload (ns = 2) so that `vs_sample_partition` compares like with like;
    # the traces for the parity check come from one more run per exchange path with ns = 6 (nt // 5 steps apart)
    Wt = torch.zeros((nsamp, a.ns), dtype=torch.float64, device="cuda")
    Wd = torch.zeros((nsamp, ns), dtype=torch.float64, device="cuda"); Kd = torch.zeros_like(Wd); Md = torch.zeros_like(Wd)

    def go():
        hc.run(chis, 0.1, nt, a.ns, W=Wt)
    res = {}
    traces = {}
    for mode in (("peer", "1"), ("nccl", "0")):
        if mode[0] == "peer" and not peer:
            continue
        os.environ["RBM_PEER_XCH"] = mode[1]
        go(); go()                                      # direct, then graph capture
        l0 = cctx.launch_count()
        ms = timed_runs(torch, go, 3, dist)
        res[mode[0]] = {"ms_per_run": ms, "value": nsamp * a.n * nt / (ms * 1e-3), "unit": UNIT,
                        "gpu_launches": int(cctx.launch_count() - l0)}
        hc.run(chis, 0.1, nt, ns, W=Wd, K=Kd, M=Md)
        torch.cuda.synchronize()
        traces[mode[0]] = (Wd.cpu().numpy().copy(), Kd.cpu().numpy().copy(), Md.cpu().numpy().copy())
    os.environ.pop("RBM_PEER_XCH", None)
    hc.close()
    # single-rank run of the same particles (every rank does it; they all hold the full particle set anyway)
    solo = r.Context(local)
    solo.set_stream(torch.cuda.current_stream().cuda_stream)
    hs = r.PICHandle(a.n, poisson, solo)
    hs.set_particles(torch.from_numpy(x).cuda(), torch.from_numpy(v).cuda(), wu)
    Ws = torch.zeros((nsamp, ns), dtype=torch.float64, device="cuda"); Ks = torch.zeros_like(Ws); Ms = torch.zeros_like(Ws)
    hs.run(chis, 0.1, nt, ns, W=Ws, K=Ks, M=Ms)
    torch.cuda.synchronize()
    errs = {}
    refs = [t.cpu().numpy() for t in (Ws, Ks, Ms)]
    for name, ref, q in zip("WKM", refs, range(3)):
        errs[name] = max(float(np.max(np.abs(tr[q] - ref)) / np.max(np.abs(ref))) for tr in traces.values())
    hs.close(); solo.close()
    worst = torch.tensor([max(errs.values())], dtype=torch.float64, device="cuda")
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    tol = 1e-11
    best = res.get("peer", res.get("nccl"))
    out["cfg2_particle_chunks"] = {
        "workload": f"BASELINE configs[2]: {nsamp} samples x {a.n:.0e} particles, nh={a.nh}, nt={nt}, ns={a.ns}; particle chunks of "
                    f"{hi - lo} over {world} ranks, charge density ({nsamp} x {a.nh} doubles) summed over ranks every step",
        "collective": {"peer": "fused NVLink push exchange inside k_step (peer memory, no collective call)" if peer else None,
                       "nccl": "ncclAllReduce between the step kernels (RBM_PEER_XCH=0)"},
        "peer": res.get("peer"), "nccl": res.get("nccl"),
        "value": best["value"], "unit": UNIT,
        "vs_sample_partition": best["value"] / value_sample_partition,
        "parity": {"against": f"single-rank run of the same particles and samples (same library, no communicator), ns={ns} saved "
                              "steps, both exchange paths",
                   "max_rel_err": {k: errs[k] for k in "WKM"}, "max_over_ranks": float(worst.item()), "tol": tol,
                   "ok": bool(float(worst.item()) <= tol)}}
    cctx_keep = cctx

    # ---- (b) BASELINE configs[3]: row-sharded Gram of a (2e7 x 2000) snapshot matrix + one NCCL all-reduce
    cols = 2000
    rows = min(a.gram_rows // world, 5_000_000)          # cap by HBM: 5e6 x 2000 x 8 B = 80 GB per rank
    free = torch.cuda.mem_get_info()[0]
    while rows * cols * 8 * 1.15 > free and rows > 100_000:
        rows //= 2
    m = cols // 2
    torch.manual_seed(1234 + rank)
    X = torch.empty((m, rows), dtype=torch.float64, device="cuda"); V = torch.empty_like(X)
    for T in (X, V):                                     # filled in slabs: randn of 40 GB at once needs a same-size temporary
        for c0 in range(0, m, 50):
            T[c0:c0 + 50].normal_()
    X = X.T; V = V.T
    G = torch.empty(cols, cols, dtype=torch.float64, device="cuda")
    keep, P, N, Ld, nb, nrows, C = r.reducedbasis._blocks((X, V))

    def gram():
        cctx_keep.check(cctx_keep.lib.rbm_gram(cctx_keep.h, P, N, Ld, nb, nrows, G.data_ptr()))
    gram()
    ms = timed_runs(torch, gram, 2, dist)
    # cheap global check: trace(G) = sum of the squared column norms over all ranks; G symmetric
    tr_local = torch.zeros(1, dtype=torch.float64, device="cuda")
    for T in (X, V):
        for c0 in range(0, m, 50):
            blk = T[:, c0:c0 + 50]
            tr_local += (blk * blk).sum()
    dist.all_reduce(tr_local)
    tr_err = abs(float(torch.diagonal(G).sum().item()) - float(tr_local.item())) / float(tr_local.item())
    sym = bool(torch.equal(G, G.T))
    R = rows * world
    flops = R * cols * (cols + 1)
    out["cfg3_distributed_gram"] = {
        "workload": f"BASELINE configs[3]: Gram of a {R} x {cols} snapshot matrix [X V], {rows} rows per rank on {world} ranks"
                    + ("" if R == a.gram_rows else f" (the full {a.gram_rows} rows need >= 4 ranks of 180 GB)"),
        "collective": "one ncclAllReduce of the 2000 x 2000 Gram (32 MB)",
        "ms": ms, "tflops_aggregate": flops / (ms * 1e-3) / 1e12, "snapshot_gbs_aggregate": 8 * R * cols / (ms * 1e-3) / 1e9,
        "tflops_per_gpu": flops / (ms * 1e-3) / 1e12 / world,
        "roofline": {"bound": "tensor", "unit": "TFLOP/s", "achieved": flops / (ms * 1e-3) / 1e12 / world, "peak": dgemm_tf,
                     "frac": flops / (ms * 1e-3) / 1e12 / world / dgemm_tf, "note": "per GPU, all-reduce inside the timed region; "
                     "peak = cuBLAS DGEMM measured in this process"},
        "check": {"trace_rel_err": tr_err, "symmetric": sym, "ok": bool(tr_err < 1e-12 and sym)}}
    del X, V, G
    torch.cuda.empty_cache()
    cctx.close()
    return out


def run_b200(a):
    import torch
    import torch.distributed as dist

    import rbm_b200 as r

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    ctx = r.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)

    L = r.BumpOnTail.domain_length(KAPPA)
    x, v, w = r.BumpOnTail.draw_accept_reject(a.n, seed=1234)
    wu = float(w[0])
    # one parameter sample per GPU: chi_r from the training range (1.0 on a single GPU)
    chi = np.array([1.0 if world == 1 else 1 / 3 + (4 / 3) * rank / (world - 1)])
    poisson = r.PoissonSolverPBSplines(3, a.nh, L)
    h = r.PICHandle(a.n, poisson, ctx)

    # ---- value: everything resident in HBM
    xd = torch.from_numpy(x).cuda(); vd = torch.from_numpy(v).cuda()
    Xd = torch.empty((1, a.ns, a.n), dtype=torch.float64, device="cuda")
    Vd = torch.empty_like(Xd); Ad = torch.empty_like(Xd)
    Wd = torch.empty((1, a.ns), dtype=torch.float64, device="cuda"); Kd = torch.empty_like(Wd); Md = torch.empty_like(Wd)
    h.set_particles(xd, vd, wu)

    def step_device(nt=None):
        h.run(chi, 0.1, nt or a.nt, a.ns, X=Xd, V=Vd, A=Ad, W=Wd, K=Kd, M=Md)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(1.0)                 # let nvidia-smi finish its start-up before anything is timed
    for _ in range(a.warmup):
        step_device()
    barrier()
    t_begin = time.perf_counter()
    l0 = ctx.launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step_device()
    e1.record()
    barrier()
    t_end = time.perf_counter()
    launches = ctx.launch_count() - l0
    ms_local = e0.elapsed_time(e1)
    ms = torch.tensor([ms_local], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * a.steps * a.nt * a.n / (ms_total * 1e-3)
    W_final = float(Wd[0, -1].item())
    ms_step_local = ms_local / a.steps          # this rank's own bench step (the roofline below is this rank's kernel)

    # ---- roofline of the dominant kernel (k_step).  Production run = one graph launch per bench step with programmatic
    # dependent launch between the step kernels, where a kernel has no individual duration; its time is the timed bench step
    # minus the other kernels of the step, which are event-timed one by one in a serialised profile pass.
    ctx.profile_enable(True)
    step_device()
    prof = {k: ctx.profile_query(k) for k in ("k_step", "k_step_save", "k_step_dual", "k_deposit", "k_solve", "k_broadcast", "k_gather")}
    ctx.profile_enable(False)
    n_step, ms_step_ser = prof["k_step"]
    other_ms = sum(ms_k for k, (_, ms_k) in prof.items() if k != "k_step")
    avg_prod = max(ms_step_local - other_ms, 0.0) / max(n_step, 1)
    avg_ser = ms_step_ser / max(n_step, 1)
    # cross-check by the slope over nt: (T(nt) - T(nt/2)) / (nt - nt/2) in the production mode; all fixed parts cancel
    nt2 = max(2, a.nt // 2)
    slope_ms = None
    if nt2 < a.nt:
        for _ in range(3):
            step_device(nt2)
        t_half = timed_runs(torch, lambda: step_device(nt2), 5)
        for _ in range(3):
            step_device()
        t_full = timed_runs(torch, step_device, 5)
        slope_ms = (t_full - t_half) / (a.nt - nt2)
    peak, peak_src = measured_peak()
    achieved = 32.0 * a.n / (avg_prod * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("k_step_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_step<SAVE=0,WEIGHTED=0>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": 32.0 * a.n, "avg_launch_ms": avg_prod, "launches_timed": n_step,
                "avg_launch_ms_serialised": avg_ser, "frac_serialised": 32.0 * a.n / (avg_ser * 1e-3) / 1e9 / peak,
                "avg_launch_ms_slope": slope_ms,
                "share_of_step": {"k_step": avg_prod * n_step, "other_kernels": other_ms, "bench_step": ms_step_local,
                                  "other_kernels_detail": {k: ms_k for k, (_, ms_k) in prof.items() if k != "k_step" and ms_k > 0},
                                  "unit": "ms per bench step"},
                "note": "avg_launch_ms = (timed bench step of the production run [graph replay + programmatic dependent launch] "
                        "- event-timed durations of the step's other kernels) / k_step launches; avg_launch_ms_serialised = "
                        "CUDA-event pair around every launch in a separate pass with PDL and graph replay off; "
                        "avg_launch_ms_slope = (T(nt) - T(nt/2)) / (nt/2) of production runs; "
                        "frac_in_step = algorithmic bytes of ALL kernels of a bench step / timed step",
                "frac_in_step": (32.0 * a.n * a.nt + 24.0 * a.n * a.ns) / (ms_step_local * 1e-3) / 1e9 / peak}

    # ---- e2e: the same call with host (pinned) buffers
    xh = torch.from_numpy(x).pin_memory(); vh = torch.from_numpy(v).pin_memory()
    Xh = torch.empty((1, a.ns, a.n), dtype=torch.float64).pin_memory()
    Vh = torch.empty_like(Xh).pin_memory(); Ah = torch.empty_like(Xh).pin_memory()
    Wh = torch.empty((1, a.ns), dtype=torch.float64).pin_memory(); Kh = torch.empty_like(Wh).pin_memory()
    Mh = torch.empty_like(Wh).pin_memory()

    def step_host():
        h.set_particles(xh, vh, wu)                                  # H2D of this step's inputs
        h.run(chi, 0.1, a.nt, a.ns, X=Xh, V=Vh, A=Ah, W=Wh, K=Kh, M=Mh)   # D2H of the snapshots and traces
        return float(Wh[0, -1])

    for _ in range(min(a.warmup, 2)):
        step_host()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(a.steps):
        W_host = step_host()
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.summary(t_begin, time.perf_counter())   # `value` region through the end of the e2e region
    clocks["window"] = "from the start of the `value` timed region to the end of the `e2e` timed region, 200 ms period"
    sampler.stop()      # every nvidia-smi query briefly stalls the GPU's work: the side measurements below run without it
    ms2 = torch.tensor([max(e0.elapsed_time(e1), 0.0)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms = float(ms2.item())
    e2e_val = world * a.steps * a.nt * a.n / (e2e_ms * 1e-3)
    h2d = 2 * 8 * a.n
    d2h = 3 * 8 * a.n * a.ns + 3 * 8 * a.ns
    assert abs(W_host - W_final) <= 1e-9 * abs(W_final), "host and device runs disagree"
    del xh, vh, Xh, Vh, Ah

    out = None
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "np": a.n, "nh": a.nh, "nt": a.nt, "ns": a.ns,
                       "samples_per_gpu": 1, "partition": "parameter samples across ranks, no data-path collective "
                       "(the communicating paths are under `multi`)",
                       "l2": "inputs larger than L2 (160 MB particle state per sample, streamed every PIC step)",
                       "particle_steps_per_bench_step": a.nt * a.n, "W_final": W_final},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / a.steps, "wall_ms_per_step": wall_ms / a.steps, "host_buffers": "pinned (torch)"},
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
    # ---- side measurements on rank 0 at N = 1 only
    dgemm_tf = None
    failed = None
    if world == 1:
        if not a.no_extras:
            try:
                out["save_every_step"] = save_every_step_measurement(a, h, chi, torch, peak)
            except Exception as e:                     # a side measurement must never cost the headline line
                out["save_every_step"] = {"error": str(e)}
            try:
                var = host_buffer_variants(a, h, ctx, chi, x, v, wu, torch, W_final)
                out["e2e"]["pageable"] = var["pageable"]; out["e2e"]["registered"] = var["registered"]
                out["e2e"]["trainingset"] = var["trainingset"]
            except Exception as e:
                out["e2e"]["variants_error"] = str(e)
        if not a.no_cpu:
            nthreads = host_threads()
            rate, nt_s, secs = cpu_oracle_rate(x, v, wu, L, a.nh, a.cpu_seconds, nthreads)
            out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": nthreads, "kind": "port",
                                   "sample": (f"same particles and grid, {nt_s} of {a.nt} time steps ({secs:.1f} s of CPU work)"
                                              + (" = the full workload" if nt_s >= a.nt else ""))}
        del xd, vd, Xd, Vd, Ad
        h.close()
        torch.cuda.empty_cache()
        if not (a.no_pod and a.no_extras):
            dgemm_tf = cublas_dgemm_tflops(torch)
        if not a.no_pod:
            try:
                out["pod"] = pod_side_measurement(ctx, torch, dgemm_tf)
            except Exception as e:
                out["pod"] = {"error": str(e)}
        if not a.no_extras:
            try:
                out["reduced_sweep"] = reduced_sweep_measurement(ctx, torch, dgemm_tf)
            except Exception as e:
                out["reduced_sweep"] = {"error": str(e)}
    else:
        del xd, vd, Xd, Vd, Ad
        h.close()
        torch.cuda.empty_cache()
        if not a.no_multi:
            try:
                dgemm_tf = cublas_dgemm_tflops(torch)
                multi = multi_gpu_section(a, torch, dist, r, x, v, wu, L, value, dgemm_tf)
                if rank == 0:
                    out["multi"] = multi
                    if not multi["cfg2_particle_chunks"]["parity"]["ok"] or not multi["cfg3_distributed_gram"]["check"]["ok"]:
                        failed = "multi-GPU parity check failed (see multi.*.parity / check)"
            except Exception as e:
                if rank == 0:
                    out["multi"] = {"error": f"{type(e).__name__}: {e}"}
    sampler.stop()
    if rank == 0:
        print(json.dumps(out))
        sys.stdout.flush()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    if failed:
        raise SystemExit(failed)


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
