#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native hot path (contract: see the task brief / DESIGN.md).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference path (oracle port)

Workload = BASELINE.json configs[1]: bump-on-tail 1D1V PIC, np = 1e7 particles, cubic B-splines with nh = 64
modes, ONE parameter sample per GPU, dt = 0.1, nt = 250 steps, ns = 2 saved steps (initial + final).
One bench "step" = one complete rbm_pic_run of that workload (nt PIC time steps).  Metric = particle-steps/s
(FP64) = n_gpus * nt * np / time.  With N > 1 (torchrun, one rank per GPU) the samples are partitioned across
ranks (independent objects, no data-path collective): weak scaling.

  value      inputs and outputs resident in HBM (device pointers through the C ABI), CUDA events, max over ranks
  e2e        the same call with HOST (pinned) buffers: H2D of x, v and D2H of X, V, A, W, K, M inside the timed region
  roofline   k_step: 32 B per particle-step / its mean launch duration (CUDA-event pairs on the launching stream)
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
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
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

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
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


def pod_side_measurement(ctx, torch, rows=1_000_000, cols=2000):
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
    # cuBLAS DGEMM denominator (not in MEASURED_PEAKS.json): 8192^3, best of 5
    Aq = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(Aq, Aq); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    dgemm_tf = 2 * 8192 ** 3 / best / 1e12
    flops = rows * cols * (cols + 1)
    tiles = (cols + 127) // 128
    del X, V, G, Aq
    torch.cuda.empty_cache()
    return {"workload": f"Gram of [X V], {rows} rows x {cols} columns (two column blocks), FP64 DMMA",
            "snapshot_gbs": 8 * rows * cols / wall / 1e9, "tflops": flops / wall / 1e12,
            "kernel_tflops": flops / kern / 1e12, "ms": wall * 1e3, "kernel_ms": kern * 1e3,
            "cublas_dgemm_tflops_measured": dgemm_tf, "frac_of_measured_dgemm": flops / kern / 1e12 / dgemm_tf,
            # what the tensor pipe actually executes: whole 128 x 128 tiles of the lower triangle (padding + full
            # diagonal tiles), against the DMMA ceiling measured by tools/microbench/dmma_occupancy.cu on this GPU type
            "executed_tflops": 2.0 * rows * (tiles * (tiles + 1) // 2) * 128 * 128 / kern / 1e12,
            "dmma_peak_tflops_microbench": 37.1,
            "roofline": {"bound": "tensor", "unit": "TFLOP/s", "achieved": flops / kern / 1e12, "peak": dgemm_tf,
                         "frac": flops / kern / 1e12 / dgemm_tf,
                         "note": "FP64 has no entry in MEASURED_PEAKS.json: peak = cuBLAS DGEMM measured in this process"}}


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

    def step_device():
        h.run(chi, 0.1, a.nt, a.ns, X=Xd, V=Vd, A=Ad, W=Wd, K=Kd, M=Md)

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
    clocks = sampler.summary(t_begin, t_end)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * a.steps * a.nt * a.n / (ms_total * 1e-3)
    W_final = float(Wd[0, -1].item())

    # ---- roofline of the dominant kernel (k_step): mean launch duration from CUDA-event pairs on the stream
    ctx.profile_enable(True)
    step_device()
    n_step, ms_step = ctx.profile_query("k_step")
    n_solve, ms_solve = ctx.profile_query("k_solve")
    n_sv, ms_sv = ctx.profile_query("k_step_save")
    n_dep, ms_dep = ctx.profile_query("k_deposit")
    ctx.profile_enable(False)
    peak, peak_src = measured_peak()
    avg_ms = ms_step / max(n_step, 1)
    achieved = 32.0 * a.n / (avg_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("k_step_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_step<SAVE=0,WEIGHTED=0>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": 32.0 * a.n, "avg_launch_ms": avg_ms, "launches_timed": n_step,
                "share_of_step": {"k_step": ms_step, "k_step_save": ms_sv, "k_solve": ms_solve, "k_deposit": ms_dep,
                                  "unit": "ms per bench step"},
                "note": "avg_launch_ms is event-timed per launch in a separate profile pass, where programmatic dependent "
                        "launch and graph replay are off (back-to-back kernels overlap otherwise and have no individual "
                        "duration); frac_in_step = algorithmic bytes of ALL kernels of a bench step / timed step",
                "frac_in_step": (32.0 * a.n * a.nt) / (ms_total / a.steps * 1e-3) / 1e9 / peak}

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
    ms2 = torch.tensor([max(e0.elapsed_time(e1), 0.0)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms = float(ms2.item())
    e2e_val = world * a.steps * a.nt * a.n / (e2e_ms * 1e-3)
    h2d = 2 * 8 * a.n
    d2h = 3 * 8 * a.n * a.ns + 3 * 8 * a.ns
    assert abs(W_host - W_final) <= 1e-9 * abs(W_final), "host and device runs disagree"

    out = None
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "np": a.n, "nh": a.nh, "nt": a.nt, "ns": a.ns,
                       "samples_per_gpu": 1, "partition": "parameter samples across ranks, no data-path collective",
                       "l2": "inputs larger than L2 (160 MB particle state per sample, streamed every PIC step)",
                       "particle_steps_per_bench_step": a.nt * a.n, "W_final": W_final},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / a.steps, "wall_ms_per_step": wall_ms / a.steps},
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
    # ---- side measurements on rank 0 at N = 1 only
    if rank == 0 and world == 1:
        if not a.no_cpu:
            from oracle import pic_oracle as po

            nthreads = host_threads()
            rate, nt_s, secs = cpu_oracle_rate(x, v, wu, L, a.nh, a.cpu_seconds, nthreads)
            out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": nthreads, "kind": "port",
                                   "sample": (f"same particles and grid, {nt_s} of {a.nt} time steps ({secs:.1f} s of CPU work)"
                                              + (" = the full workload" if nt_s >= a.nt else ""))}
        if not a.no_pod:
            del xd, vd, Xd, Vd, Ad
            torch.cuda.empty_cache()
            try:
                out["pod"] = pod_side_measurement(ctx, torch)
            except Exception as e:                     # the side measurement must never cost the headline line
                out["pod"] = {"error": str(e)}
    sampler.stop()
    if rank == 0:
        print(json.dumps(out))
    h.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
