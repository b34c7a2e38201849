"""GPU parity tests of the full-order PIC path, through the C ABI (ctypes), against the CPU oracle.

Bars (BASELINE.json north_star): particle cell indices bit-exact; charge density and fields <= 1e-12
relative; energy/momentum traces <= 1e-10.  "Relative" is to the largest magnitude of the compared
array.  The oracle accumulates the deposit in extended precision so that the tolerance measures the
GPU's error, not the oracle's own summation roundoff.
"""
import os

import numpy as np
import pytest

from conftest import rel
from oracle import pic_oracle as po

pytestmark = pytest.mark.gpu

import rbm_b200 as r  # noqa: E402

L = r.BumpOnTail.domain_length(0.3)
TOL_FIELD = 1e-12
TOL_TRACE = 1e-10


def handle(ctx, n, nh):
    return r.PICHandle(n, r.PoissonSolverPBSplines(3, nh, L), ctx)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "pic_small.npz"))


# ------------------------------------------------------------------ cell index: bit-exact
@pytest.mark.parametrize("nh", [16, 64, 100, 1000])
def test_locate_bit_exact(ctx, nh):
    rng = np.random.default_rng(nh)
    h = L / nh
    knots = np.arange(nh + 1) * h
    x = np.concatenate([
        rng.random(200_000) * L, -rng.random(50_000) * 30 * L, rng.random(50_000) * 30 * L,
        knots, np.nextafter(knots, -np.inf), np.nextafter(knots, np.inf), -knots, knots + 5 * L, knots - 7 * L,
        np.array([0.0, -0.0, L, -L, 2 * L, -1e-300, 1e-300, 1e-17, -1e-17, np.nextafter(L, 0), -np.nextafter(L, 0),
                  1e9, -1e9, 1e14, 4.0e15 * L, 1e300, -1e300]),
    ])
    hnd = handle(ctx, 16, nh)
    cell, xi = hnd.locate(x)
    c0, xi0 = po.locate(x, L, nh)
    assert np.array_equal(cell, c0), f"{np.count_nonzero(cell != c0)} cell indices differ"
    assert np.array_equal(xi, xi0), f"{np.count_nonzero(xi != xi0)} offsets differ"
    hnd.close()


def test_locate_golden_and_other_domain(ctx, gold):
    hnd = handle(ctx, 16, 16)
    c, xi = hnd.locate(gold["loc_x"])
    assert np.array_equal(c, gold["loc_cell"]) and np.array_equal(xi, gold["loc_xi"])
    hnd.close()
    for LL, nh in ((2 * np.pi, 10), (1.0, 7), (3.3, 100)):
        hh = r.PICHandle(8, r.PoissonSolverPBSplines(3, nh, LL), ctx)
        x = (np.random.default_rng(nh).random(100_000) - 0.3) * 5 * LL
        c, xi = hh.locate(x)
        c0, xi0 = po.locate(x, LL, nh)
        assert np.array_equal(c, c0) and np.array_equal(xi, xi0)
        hh.close()


# ------------------------------------------------------------------ deposit / solve / gather
@pytest.mark.parametrize("n,nh", [(10_000, 16), (100_003, 64), (1_000_000, 64), (50_000, 100), (20_000, 128),
                                  (20_000, 256), (5_000, 1000)])
def test_deposit_uniform_and_weighted(ctx, n, nh):
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=n % 97)
    x = x + np.floor(np.random.default_rng(1).random(n) * 7 - 3) * L      # unwrapped positions
    hnd = handle(ctx, n, nh)
    rhs = hnd.deposit(x, float(w[0]))
    ref = po.deposit(x, float(w[0]), L, nh, extended=True, nthreads=4)
    assert rel(rhs, ref) < TOL_FIELD
    assert abs(rhs.sum() - L) < 1e-12 * L
    assert rel(rhs - rhs.mean(), ref - ref.mean()) < 50 * TOL_FIELD        # after the ~30x cancellation
    wv = np.random.default_rng(2).random(n) * 2 * L / n
    assert rel(hnd.deposit(x, wv), po.deposit(x, wv, L, nh, extended=True, nthreads=4)) < TOL_FIELD
    hnd.close()


def test_deposit_edge_cases(ctx):
    hnd = handle(ctx, 8, 16)
    one = hnd.deposit(np.array([0.25 * L]), 2.0)                           # a single particle
    assert rel(one, po.deposit(np.array([0.25 * L]), 2.0, L, 16)) < 1e-15
    assert np.count_nonzero(one) <= 4
    same = hnd.deposit(np.full(100_000, 1.234), 1e-5)                      # every particle in one cell
    assert rel(same, po.deposit(np.full(100_000, 1.234), 1e-5, L, 16)) < TOL_FIELD
    with pytest.raises(r.RBMError):
        hnd.deposit(np.zeros(4), 0.0)                                      # w_uniform must be positive
    # a device array that starts 8 bytes off a 16-byte boundary takes the scalar-load path of k_deposit: bit-identical
    import torch
    xs, _, _ = r.BumpOnTail.draw_accept_reject(50_001, seed=8)
    big = torch.from_numpy(np.concatenate([[0.0], xs])).cuda()
    assert big[1:].data_ptr() % 16 == 8
    assert np.array_equal(hnd.deposit(big[1:], 1e-4), hnd.deposit(torch.from_numpy(xs).cuda(), 1e-4))
    hnd.close()


@pytest.mark.parametrize("nh,chi", [(16, 1.0), (64, 1.0 / 3.0), (64, 5.0 / 3.0), (100, 0.9)])
def test_field_protocol(ctx, nh, chi):
    """update! / efield! / energy / coefficients (src/particles/electric_field.jl:22-35)."""
    n = 200_000
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=11)
    f = r.ScaledPoissonField(r.PoissonSolverPBSplines(3, nh, L), chi, ctx)
    f.update(x, w)
    rhs = po.deposit(x, w, L, nh, extended=True, nthreads=4)
    phi0, W0 = po.poisson(rhs, nh, L, chi)
    assert rel(f.coefficients(), phi0) < TOL_FIELD
    assert abs(f.energy() - W0) < TOL_FIELD * abs(W0)
    xe = np.random.default_rng(3).random(637) * 3 * L - L                  # DEIM-style: k_e evaluation points
    E = np.zeros(637)
    f.efield(E, xe)
    assert rel(E, po.gather(phi0, xe, L, nh)) < TOL_FIELD
    E2 = np.zeros(n)
    f(E2, x, float(w[0]))                                                  # callable form f(E, x, w, t)
    assert rel(E2, po.gather(phi0, x, L, nh)) < TOL_FIELD
    with pytest.raises(ValueError):
        f.update(x, w[:-1])


# ------------------------------------------------------------------ one step from identical inputs
@pytest.mark.parametrize("n,nh,chi", [(10_000, 16, 1.0), (65_537, 64, 5.0 / 3.0), (300_000, 64, 1.0 / 3.0)])
def test_single_step_parity(ctx, n, nh, chi):
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=5)
    hnd = handle(ctx, n, nh)
    x1, v1, rhs, phi, a = hnd.step(x, v, float(w[0]), chi, 0.1)
    X1, V1, RHS, PHI, A = po.step(x, v, float(w[0]), L, nh, chi, 0.1, extended=True, nthreads=4)
    assert rel(rhs, RHS) < TOL_FIELD and rel(phi, PHI) < TOL_FIELD and rel(a, A) < TOL_FIELD
    assert np.max(np.abs(v1 - V1)) < 1e-13 and np.max(np.abs(x1 - X1)) < 1e-13
    # cell indices of the stepped particles: bit-exact wherever the positions are bit-identical
    same = x1 == X1
    assert same.mean() > 0.5
    c_gpu, _ = hnd.locate(x1[same]); c_ref, _ = po.locate(X1[same], L, nh)
    assert np.array_equal(c_gpu, c_ref)
    # the drift is never contracted into an FMA: feeding the oracle's own acceleration reproduces x1, v1 bitwise
    dte = 0.1 * chi; hdt = 0.5 * dte
    xp = x + hdt * v; vv = v + dte * a; xx = xp + hdt * vv
    assert np.array_equal(v1, vv) and np.array_equal(x1, xx)
    hnd.close()


# ------------------------------------------------------------------ trajectories
def _run(hnd, chis, dt, nt, ns, n, nh, what="XVAPWKM"):
    p = len(chis)
    out = {}
    if "X" in what: out["X"] = np.zeros((p, ns, n))
    if "V" in what: out["V"] = np.zeros((p, ns, n))
    if "A" in what: out["A"] = np.zeros((p, ns, n))
    out["Phi"] = np.zeros((p, ns, nh)); out["W"] = np.zeros((p, ns)); out["K"] = np.zeros((p, ns)); out["M"] = np.zeros((p, ns))
    hnd.run(chis, dt, nt, ns, X=out.get("X"), V=out.get("V"), A=out.get("A"), Phi=out["Phi"], W=out["W"], K=out["K"], M=out["M"])
    return out


def test_golden_trajectories(ctx, gold):
    n, nh, nt, ns, dt = 2000, int(gold["nh"]), int(gold["nt"]), int(gold["ns"]), float(gold["dt"])
    hnd = handle(ctx, n, nh)
    hnd.set_particles(gold["x0"], gold["v0"], float(gold["w"]))
    o = _run(hnd, gold["chi"], dt, nt, ns, n, nh)
    for p in range(3):
        for k in ("W", "K", "M"):
            assert rel(o[k][p], gold[f"{k}_{p}"]) < TOL_TRACE, (k, p)
        assert rel(o["Phi"][p], gold[f"Phi_{p}"]) < 1e-10
        assert rel(o["X"][p, -1], gold[f"Xlast_{p}"]) < 1e-11
        assert rel(o["V"][p, -1], gold[f"Vlast_{p}"]) < 1e-10
        assert rel(o["A"][p, -1], gold[f"Alast_{p}"]) < 1e-9
    hnd.close()


def test_shipped_defaults_config1(ctx):
    """scripts/bump_on_tail_1_training.jl defaults: np=1e4, nh=16, p=3, dt=0.1, nt=250, ns=251, chi=LinRange(1/3,5/3,10)."""
    n, nh, nt, ns, dt = 10_000, 16, 250, 251, 0.1
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=1234)
    chis = np.linspace(1 / 3, 5 / 3, 10)
    hnd = handle(ctx, n, nh)
    hnd.set_particles(x, v, float(w[0]))
    o = _run(hnd, chis, dt, nt, ns, n, nh)
    worst = {}
    for p, chi in enumerate(chis):
        ref = po.integrate(x, v, float(w[0]), L, nh, float(chi), dt, nt, ns, extended=True, nthreads=4)
        for k in ("W", "K", "M"):
            worst[k] = max(worst.get(k, 0.0), rel(o[k][p], ref[k]))
        worst["Phi"] = max(worst.get("Phi", 0.0), rel(o["Phi"][p], ref["Phi"]))
        worst["X"] = max(worst.get("X", 0.0), rel(o["X"][p], ref["X"]))
        # saved slot 0 is the initial condition itself
        assert np.array_equal(o["X"][p, 0], x) and np.array_equal(o["V"][p, 0], v)
    print("config1 worst relative deviations:", worst)
    assert worst["W"] < TOL_TRACE and worst["K"] < TOL_TRACE and worst["M"] < TOL_TRACE
    assert worst["Phi"] < 1e-8 and worst["X"] < 1e-9
    E = o["K"] + o["W"]
    assert np.all((E.max(axis=1) - E.min(axis=1)) / E[:, 0] < 5e-4)
    hnd.close()


def test_thinned_saving_device_outputs_and_determinism(ctx):
    import torch

    n, nh, nt, dt = 50_001, 64, 24, 0.1           # odd np: scalar snapshot stores, unaligned slots
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=9)
    chis = np.array([0.5, 1.0, 1.5])
    hnd = handle(ctx, n, nh)
    hnd.set_particles(x, v, float(w[0]))
    full = _run(hnd, chis, dt, nt, nt + 1, n, nh)
    thin = _run(hnd, chis, dt, nt, 5, n, nh)                                 # nsave = 6
    # The save cadence changes which steps run the saved-step kernel, whose deposit sums the same numbers in another
    # (fixed) order: trajectories agree to round-off between cadences, and bit for bit for a given cadence.
    for k in ("X", "V"):
        assert np.max(np.abs(thin[k] - full[k][:, [0, 6, 12, 18, 24]])) < 1e-12, k
    for k in ("A", "Phi", "W", "K", "M"):
        assert rel(thin[k], full[k][:, [0, 6, 12, 18, 24]]) < 1e-12, k
    again = _run(hnd, chis, dt, nt, 5, n, nh)
    for k in thin:
        assert np.array_equal(thin[k], again[k]), k                          # run-to-run bit reproducible
    # device-resident outputs (what the POD stage consumes without a host round trip)
    Xd = torch.zeros((3, 5, n), dtype=torch.float64, device="cuda")
    Wd = torch.zeros((3, 5), dtype=torch.float64, device="cuda")
    hnd.run(chis, dt, nt, 5, X=Xd, W=Wd)
    assert np.array_equal(Xd.cpu().numpy(), thin["X"]) and np.array_equal(Wd.cpu().numpy(), thin["W"])
    # sample independence: one sample alone gives the same bits as inside the batch
    solo = _run(hnd, chis[1:2], dt, nt, 5, n, nh)
    assert np.array_equal(solo["X"][0], thin["X"][1]) and np.array_equal(solo["W"][0], thin["W"][1])
    ref = po.integrate(x, v, float(w[0]), L, nh, 1.0, dt, nt, 5, extended=True, nthreads=4)
    assert rel(thin["W"][1], ref["W"]) < TOL_TRACE and rel(thin["X"][1], ref["X"]) < 1e-12
    hnd.close()


def test_weighted_particles_and_large_nh(ctx):
    n, nt, dt = 20_000, 10, 0.1
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=4)
    wv = w * (0.5 + np.random.default_rng(0).random(n))
    for nh in (16, 64, 1000):                                                # 1000: global-atomic fallback path
        hnd = handle(ctx, n, nh)
        hnd.set_particles(x, v, wv)
        o = _run(hnd, np.array([1.0]), dt, nt, nt + 1, n, nh)
        ref = po.integrate(x, v, wv, L, nh, 1.0, dt, nt, nt + 1, extended=True, nthreads=4)
        for k in ("W", "K", "M"):
            assert rel(o[k][0], ref[k]) < TOL_TRACE, (nh, k)
        assert rel(o["X"][0], ref["X"]) < 1e-12 and rel(o["A"][0], ref["A"]) < 1e-10
        hnd.close()


def test_error_behaviour(ctx):
    """The reference signals bad arguments with @assert / DimensionMismatch; the C ABI returns codes."""
    with pytest.raises(r.RBMError) as e:
        r.PICHandle(10, r.PoissonSolverPBSplines(4, 16, L), ctx)            # degrees 1, 2, 3 are implemented
    assert e.value.code == -4
    with pytest.raises(r.RBMError):
        r.PICHandle(10, r.PoissonSolverPBSplines(3, 3, L), ctx)
    with pytest.raises(r.RBMError):
        r.PICHandle(0, r.PoissonSolverPBSplines(3, 16, L), ctx)
    hnd = handle(ctx, 100, 16)
    with pytest.raises(r.RBMError) as e:
        hnd.run([1.0], 0.1, 10, 11)                                          # no particles yet
    assert e.value.code == -1 and "set_particles" in str(e.value)
    x, v, w = r.BumpOnTail.draw_accept_reject(100, seed=1)
    hnd.set_particles(x, v, float(w[0]))
    for bad in (dict(chi=[1.0], nt=10, ns=1), dict(chi=[1.0], nt=3, ns=11), dict(chi=[-1.0], nt=10, ns=11),
                dict(chi=[], nt=10, ns=11)):
        with pytest.raises(r.RBMError):
            hnd.run(bad["chi"], 0.1, bad["nt"], bad["ns"])
    with pytest.raises(r.RBMError):
        hnd.field_energy()                                                   # update! not called yet
    hnd.close()


def test_integrate_vp_drop_in_and_trainingset(ctx):
    """The script loop of scripts/bump_on_tail_1_training.jl:87-109 with the upstream per-sample signature,
    and the batched variant writing straight into Snapshots."""
    n, nh, nt = 5000, 16, 20
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=8)
    particles = r.ParticleList(x[None, :], v[None, :], w[None, :])
    poisson = r.PoissonSolverPBSplines(3, nh, L)
    pspace = r.ParameterSpace(r.Parameter("chi", 1 / 3, 5 / 3, 4))
    IP = r.VPIntegratorParameters(0.1, nt, nt + 1, nh, n)
    IC = r.VPIntegratorCache(IP)
    TS = r.TrainingSet.create(particles, poisson, nt + 1, {"kappa": 0.3}, pspace, IP.with_paramspace(pspace))
    SS = TS.snapshots
    for p in pspace.eachindex():
        lparams = pspace(p)
        efield = r.ScaledPoissonField(poisson, lparams["chi"], ctx)
        r.integrate_vp(particles, efield, lparams, IP, IC, save=True, ctx=ctx)
        SS.X[p, :, :, 0] = IC.X; SS.V[p, :, :, 0] = IC.V; SS.A[p, :, :, 0] = IC.A; SS.Phi[p, :, :, 0] = IC.Phi
        SS.W[p] = IC.W; SS.K[p] = IC.K; SS.M[p] = IC.M
    assert np.array_equal(particles.x[0], x)                                 # `particles` left untouched
    batched = r.integrate_vp_all(particles, poisson, pspace.column("chi"), TS.integrator, ctx=ctx)
    assert batched == SS                                                     # bit-identical to the per-sample loop
    ref = po.integrate(x, v, float(w[0]), L, nh, float(pspace(2)["chi"]), 0.1, nt, nt + 1, extended=True)
    assert rel(SS.W[2], ref["W"]) < TOL_TRACE and rel(SS.X[2, :, :, 0], ref["X"]) < 1e-12


# ------------------------------------------------------------------ full size (BASELINE config 2): properties
def test_full_size_properties(ctx):
    """np = 1e7, nh = 64: partition of unity, a few steps against the threaded oracle, momentum bookkeeping."""
    n, nh, nt, dt = 10_000_000, 64, 4, 0.1
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=1234)
    hnd = handle(ctx, n, nh)
    rhs = hnd.deposit(x, float(w[0]))
    assert abs(rhs.sum() - L) < 1e-12 * L
    ref_rhs = po.deposit(x, float(w[0]), L, nh, extended=True, nthreads=po.max_threads())
    assert rel(rhs, ref_rhs) < TOL_FIELD
    hnd.set_particles(x, v, float(w[0]))
    W = np.zeros((1, 2)); K = np.zeros((1, 2)); M = np.zeros((1, 2)); X = np.zeros((1, 2, n))
    hnd.run([1.0], dt, nt, 2, X=X, W=W, K=K, M=M)
    ref = po.integrate(x, v, float(w[0]), L, nh, 1.0, dt, nt, 2, extended=True, nthreads=po.max_threads(), outputs="X")
    assert rel(W[0], ref["W"]) < TOL_TRACE and rel(K[0], ref["K"]) < TOL_TRACE and rel(M[0], ref["M"]) < TOL_TRACE
    same = X[0, 1] == ref["X"][1]
    assert np.max(np.abs(X[0, 1] - ref["X"][1])) < 1e-12 and same.mean() > 0.5
    c_gpu, _ = hnd.locate(X[0, 1][same][:1_000_000]); c_ref, _ = po.locate(ref["X"][1][same][:1_000_000], L, nh)
    assert np.array_equal(c_gpu, c_ref)
    hnd.close()


# ------------------------------------------------------------------ edge cases: sizes, grids, kernel variants
@pytest.mark.parametrize("n,nh", [(1, 16), (3, 4), (1537, 4), (4099, 32), (9001, 65), (7000, 128), (5000, 200),
                                  (3001, 256), (3001, 257), (2000, 1024)])
def test_grid_and_size_edges(ctx, n, nh):
    """Every (T, REP) instantiation (nh <= 29: 512; <= 64: 384; <= 128: 192; <= 256: 96), the global-atomic fallback
    (nh > 256), the minimum and maximum grid, particle counts that are not multiples of the vector width."""
    rng = np.random.default_rng(n + nh)
    x = (rng.random(n) * 3 - 1) * L
    v = rng.standard_normal(n) * 2
    w = L / n
    hnd = handle(ctx, n, nh)
    nt = 6
    hnd.set_particles(x, v, w)
    o = _run(hnd, np.array([0.8, 1.2]), 0.1, nt, nt + 1, n, nh)
    for p, chi in enumerate((0.8, 1.2)):
        ref = po.integrate(x, v, w, L, nh, chi, 0.1, nt, nt + 1, extended=True)
        assert rel(o["X"][p], ref["X"]) < 1e-12 and rel(o["V"][p], ref["V"]) < 1e-11
        assert rel(o["Phi"][p], ref["Phi"]) < 1e-10 and rel(o["K"][p], ref["K"]) < 1e-11
        assert np.max(np.abs(o["W"][p] - ref["W"])) < 1e-10 * max(np.max(np.abs(ref["W"])), 1e-300)
        assert np.max(np.abs(o["A"][p] - ref["A"])) < 1e-9 * max(np.max(np.abs(ref["A"])), 1e-300)
    hnd.close()


def test_empty_and_degenerate_inputs(ctx):
    hnd = handle(ctx, 8, 16)
    assert np.array_equal(hnd.deposit(np.zeros(0), 1.0, n=0), np.zeros(16))       # empty input
    c, xi = hnd.locate(np.zeros(0), n=0)
    assert c.shape == (0,)
    x = np.zeros(8); v = np.zeros(8)                                               # all particles at rest at the origin
    hnd.set_particles(x, v, L / 8)
    o = _run(hnd, np.array([1.0]), 0.1, 3, 4, 8, 16)
    ref = po.integrate(x, v, L / 8, L, 16, 1.0, 0.1, 3, 4, extended=True)
    assert rel(o["Phi"][0], ref["Phi"]) < 1e-11 and np.max(np.abs(o["K"][0] - ref["K"])) < 1e-25
    assert np.max(np.abs(o["X"][0] - ref["X"])) < 1e-13
    hnd.close()


def test_maximum_grid(ctx):
    """nh = 4096 (the largest grid the C ABI accepts): one step against the oracle."""
    n, nh = 3000, 4096
    rng = np.random.default_rng(1)
    x = rng.random(n) * L; v = rng.standard_normal(n)
    hnd = handle(ctx, n, nh)
    x1, v1, rhs, phi, a = hnd.step(x, v, L / n, 1.1, 0.1)
    X1, V1, RHS, PHI, A = po.step(x, v, L / n, L, nh, 1.1, 0.1, extended=True)
    assert rel(rhs, RHS) < 1e-12 and rel(phi, PHI) < 1e-9 and rel(a, A) < 1e-8      # cond(K_s) ~ (nh/pi)^2 = 1.7e6
    assert np.max(np.abs(x1 - X1)) < 1e-9
    with pytest.raises(r.RBMError):
        r.PICHandle(10, r.PoissonSolverPBSplines(3, 4097, L), ctx)
    hnd.close()


def test_graph_replay_matches_direct_launches(ctx):
    """A repeated rbm_pic_run with identical shapes and device pointers is replayed as ONE CUDA graph from the
    second repetition on (bench.py's timed loop); the replay must give the bits of the direct launches, also after
    the inputs (chi, particles) changed behind the same pointers."""
    import torch

    n, nh, nt, ns = 40_001, 64, 30, 4
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=12)
    hnd = handle(ctx, n, nh)
    xd = torch.from_numpy(x).cuda(); vd = torch.from_numpy(v).cuda()
    hnd.set_particles(xd, vd, float(w[0]))
    X = torch.zeros((2, ns, n), dtype=torch.float64, device="cuda"); V = torch.zeros_like(X); A = torch.zeros_like(X)
    W = torch.zeros((2, ns), dtype=torch.float64, device="cuda")
    outs = []
    l0 = ctx.launch_count()
    for rep in range(4):                       # 1: direct, 2: capture + replay, 3-4: replay
        X.zero_(); V.zero_(); A.zero_(); W.zero_()
        hnd.run(np.array([0.7, 1.3]), 0.1, nt, ns, X=X, V=V, A=A, W=W)
        outs.append((X.cpu().numpy().copy(), A.cpu().numpy().copy(), W.cpu().numpy().copy()))
    per_run = (ctx.launch_count() - l0) // 4
    assert (ctx.launch_count() - l0) == 4 * per_run          # replays are counted like direct launches
    for o in outs[1:]:
        for a, b in zip(o, outs[0]):
            assert np.array_equal(a, b)
    # same pointers, different parameter values and particles: the graph must pick them up
    hnd.set_particles(torch.from_numpy(x[::-1].copy()).cuda(), torch.from_numpy(v[::-1].copy()).cuda(), float(w[0]))
    hnd.run(np.array([0.9, 1.1]), 0.1, nt, ns, X=X, V=V, A=A, W=W)
    ref = po.integrate(x[::-1], v[::-1], float(w[0]), L, nh, 1.1, 0.1, nt, ns, extended=True, nthreads=4)
    assert rel(W.cpu().numpy()[1], ref["W"]) < TOL_TRACE and rel(X.cpu().numpy()[1], ref["X"]) < 1e-12
    hnd.close()


def test_host_outputs_pinned_and_pageable(ctx):
    """Staged snapshots leave the device slot by slot on a copy stream: pinned (pitched copies) and pageable
    (per-sample copies) destinations must both receive exactly what a device-resident run produces."""
    import torch

    n, nh, nt, ns = 30_000, 16, 12, 5
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=13)
    chis = np.array([0.6, 1.0, 1.4])
    hnd = handle(ctx, n, nh)
    hnd.set_particles(x, v, float(w[0]))
    dev = {k: torch.zeros((3, ns, n), dtype=torch.float64, device="cuda") for k in "XVA"}
    hnd.run(chis, 0.1, nt, ns, X=dev["X"], V=dev["V"], A=dev["A"])
    pin = {k: torch.zeros((3, ns, n), dtype=torch.float64).pin_memory() for k in "XVA"}
    hnd.run(chis, 0.1, nt, ns, X=pin["X"], V=pin["V"], A=pin["A"])
    page = {k: np.zeros((3, ns, n)) for k in "XVA"}
    hnd.run(chis, 0.1, nt, ns, X=page["X"], V=page["V"], A=page["A"])
    for k in "XVA":
        d = dev[k].cpu().numpy()
        assert np.array_equal(pin[k].numpy(), d) and np.array_equal(page[k], d), k
    only_a = np.zeros((3, ns, n))                                  # any subset of outputs may be requested
    hnd.run(chis, 0.1, nt, ns, A=only_a)
    assert np.array_equal(only_a, page["A"])
    hnd.close()


def test_host_outputs_alternating_staging_sets(ctx, monkeypatch):
    """More lock-step groups than staging sets: the two sets of staging buffers alternate between consecutive groups,
    a group's last copies run under the next group's time loop, and the group after next waits for them.  Host arrays
    (pinned and pageable) must equal a device-resident run bit for bit, also with a ragged last group and with the
    single-set path."""
    import torch

    n, nh, nt, ns = 50_000, 16, 10, 3
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=21)
    chis = np.linspace(0.5, 1.5, 7)
    hnd = handle(ctx, n, nh)
    hnd.set_particles(x, v, float(w[0]))
    monkeypatch.setenv("RBM_GROUP", "2")                            # groups of 2, 2, 2, 1 samples (same plan in every run)
    dev = {k: torch.zeros((7, ns, n), dtype=torch.float64, device="cuda") for k in "XVA"}
    Wd = np.zeros((7, ns))
    hnd.run(chis, 0.1, nt, ns, X=dev["X"], V=dev["V"], A=dev["A"], W=Wd)
    for sets2 in ("1", "0"):
        monkeypatch.setenv("RBM_STAGE_SETS2", sets2)
        pin = {k: torch.zeros((7, ns, n), dtype=torch.float64).pin_memory() for k in "XVA"}
        page = {k: np.zeros((7, ns, n)) for k in "XVA"}
        Wp = np.zeros((7, ns))
        for _ in range(2):                                          # a repeated call reuses the sets and their events
            hnd.run(chis, 0.1, nt, ns, X=pin["X"], V=pin["V"], A=pin["A"], W=Wp)
        hnd.run(chis, 0.1, nt, ns, X=page["X"], V=page["V"], A=page["A"])
        for k in "XVA":
            d = dev[k].cpu().numpy()
            assert np.array_equal(pin[k].numpy(), d), (sets2, k)
            assert np.array_equal(page[k], d), (sets2, k)
        assert np.array_equal(Wp, Wd)
    hnd.close()


def test_locate_randomised_domains_and_grids(ctx):
    """Bit-exact cell index and offset for random (L, nh) -- every one exercises a different reciprocal 1/h in the
    division-free locate -- with knot-adjacent, negative, huge and denormal positions."""
    rng = np.random.default_rng(77)
    for trial in range(24):
        LL = float(10.0 ** rng.uniform(-2, 2.5))
        nh = int(rng.integers(4, 257))
        h = LL / nh
        knots = rng.integers(-30 * nh, 30 * nh, 3000) * h
        x = np.concatenate([
            rng.uniform(-25 * LL, 25 * LL, 60_000), knots, np.nextafter(knots, np.inf), np.nextafter(knots, -np.inf),
            rng.uniform(-1, 1, 300) * 1e-300, rng.uniform(-1, 1, 300) * 1e11 * LL,
            np.array([0.0, -0.0, LL, -LL, 3 * LL, -3 * LL, np.nextafter(LL, 0), -np.nextafter(LL, 0)])])
        hh = r.PICHandle(8, r.PoissonSolverPBSplines(3, nh, LL), ctx)
        c, xi = hh.locate(x)
        c0, xi0 = po.locate(x, LL, nh)
        assert np.array_equal(c, c0), (trial, LL, nh, int(np.count_nonzero(c != c0)))
        assert np.array_equal(xi, xi0), (trial, LL, nh, int(np.count_nonzero(xi != xi0)))
        hh.close()


@pytest.mark.parametrize("n,nh,weighted", [(60_001, 64, False), (30_000, 16, True), (20_003, 128, False), (9_000, 32, False)])
def test_saved_step_second_histogram_matches_separate_deposit(ctx, n, nh, weighted):
    """Saved steps: the update! at the saved positions (time_marching.jl:95-99) rides along in k_step<DUAL> (second
    histogram).  Same Phi, W, K, M, X, A as the separate deposit pass (RBM_DUAL_SAVE=0), and as the oracle."""
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=n)
    wv = (w * np.random.default_rng(1).uniform(0.5, 1.5, n)) if weighted else float(w[0])
    nt, ns, chis = 12, 13, np.array([0.6, 1.4])
    hnd = handle(ctx, n, nh)
    hnd.set_particles(x, v, wv)
    dual = _run(hnd, chis, 0.1, nt, ns, n, nh)
    os.environ["RBM_DUAL_SAVE"] = "0"
    try:
        h2 = handle(ctx, n, nh)                 # the switch is read when the handle is created
    finally:
        del os.environ["RBM_DUAL_SAVE"]
    h2.set_particles(x, v, wv)
    sep = _run(h2, chis, 0.1, nt, ns, n, nh)
    for k in ("Phi", "W", "K", "M", "A"):
        assert rel(dual[k], sep[k]) < 1e-12, k
    assert np.max(np.abs(dual["X"] - sep["X"])) < 1e-12 and np.max(np.abs(dual["V"] - sep["V"])) < 1e-12
    ref = po.integrate(x, v, wv, L, nh, float(chis[1]), 0.1, nt, ns, extended=True, nthreads=4)
    assert rel(dual["W"][1], ref["W"]) < TOL_TRACE and rel(dual["K"][1], ref["K"]) < TOL_TRACE
    assert rel(dual["Phi"][1], ref["Phi"]) < 1e-9
    again = _run(hnd, chis, 0.1, nt, ns, n, nh)              # fixed summation order: bit-reproducible
    assert all(np.array_equal(dual[k], again[k]) for k in dual)
    hnd.close(); h2.close()


def test_registered_host_arrays(ctx):
    """rbm_host_register: pageable NumPy snapshots (what a Julia Array is) page-locked in place take the pinned copy
    path; same bytes as unregistered; double registration is an error; unregister restores the pageable path."""
    n, nh, nt, ns = 40_000, 64, 10, 6
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=5)
    hnd = handle(ctx, n, nh)
    hnd.set_particles(x, v, float(w[0]))
    chis = np.array([0.8, 1.2, 1.0])
    page = {k: np.zeros((3, ns, n)) for k in "XVA"}
    hnd.run(chis, 0.1, nt, ns, X=page["X"], V=page["V"], A=page["A"])
    reg = {k: np.zeros((3, ns, n)) for k in "XVA"}
    for a in reg.values():
        ctx.host_register(a)
    try:
        with pytest.raises(r.RBMError):
            ctx.host_register(reg["X"])
        hnd.run(chis, 0.1, nt, ns, X=reg["X"], V=reg["V"], A=reg["A"])
    finally:
        for a in reg.values():
            ctx.host_unregister(a)
    for k in "XVA":
        assert np.array_equal(reg[k], page[k]), k
    hnd.run(chis, 0.1, nt, ns, X=reg["X"])                     # pageable again
    assert np.array_equal(reg["X"], page["X"])
    hnd.close()


def test_output_buffers_are_validated(ctx):
    """The C ABI sees raw addresses; the host mirror refuses buffers the library would overrun."""
    n, nh, nt, ns = 1000, 16, 4, 5
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=2)
    hnd = handle(ctx, n, nh)
    with pytest.raises(ValueError):
        hnd.set_particles(x[:-1], v, float(w[0]))
    with pytest.raises(ValueError):
        hnd.set_particles(x.astype(np.float32), v, float(w[0]))
    hnd.set_particles(x, v, float(w[0]))
    with pytest.raises(ValueError):
        hnd.run([1.0], 0.1, nt, ns, X=np.zeros((1, ns, n), dtype=np.float32))
    with pytest.raises(ValueError):
        hnd.run([1.0], 0.1, nt, ns, X=np.zeros((1, ns - 1, n)))
    with pytest.raises(ValueError):
        hnd.run([1.0, 1.1], 0.1, nt, ns, W=np.zeros((1, ns)))
    with pytest.raises(ValueError):
        hnd.run([1.0], 0.1, nt, ns, X=np.zeros((1, ns, 2 * n))[:, :, ::2])
    ss = r.Snapshots(1, n, nh, ns + 1, 1)                      # built for another ns
    ip = r.IntegratorParameters(0.1, nt, ns, nh, n, 1)
    with pytest.raises(ValueError):
        r.integrate_vp_all(r.ParticleList(x[None], v[None], w[None]), r.PoissonSolverPBSplines(3, nh, L), [1.0], ip, ss, ctx=ctx)
    hnd.close()


@pytest.mark.parametrize("degree", [1, 2])
def test_spline_degrees_one_and_two(ctx, degree):
    """`PoissonSolverPBSplines(p, nh, L)` with p = 1, 2 (every reference script uses 3): the generic kernels with zero weights
    on the unused taps.  Deposit, solve, gather and a short run against the NumPy restatement at the same bars."""
    from oracle import pic_numpy as pn

    n, nh = 20_001, 24
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=degree)
    hnd = r.PICHandle(n, r.PoissonSolverPBSplines(degree, nh, L), ctx)
    c, xi = hnd.locate(x)
    c0, xi0 = pn.locate(x, L, nh)
    assert np.array_equal(c, c0) and np.array_equal(xi, xi0)
    rhs = pn.deposit(x, w, L, nh, degree)
    assert rel(hnd.deposit(x, float(w[0])), rhs) < TOL_FIELD
    phi, W = pn.poisson(rhs, nh, L, 0.8, degree)
    hnd.field_update(x, float(w[0]), 0.8)
    assert rel(hnd.field_coefficients(), phi) < TOL_FIELD and abs(hnd.field_energy() - W) <= TOL_TRACE * abs(W)
    assert rel(hnd.field_eval(x), pn.gather(phi, x, L, nh, degree)) < TOL_FIELD
    nt, ns = 20, 11
    hnd.set_particles(x, v, float(w[0]))
    out = _run(hnd, np.array([0.7, 1.2]), 0.1, nt, ns, n, nh)
    ref = pn.integrate(x, v, w, L, nh, 1.2, 0.1, nt, ns, degree=degree)
    for k in ("W", "K", "M"):
        assert rel(out[k][1], ref[k]) < TOL_TRACE, k
    assert np.max(np.abs(out["X"][1] - ref["X"])) < 1e-11 and rel(out["A"][1], ref["A"]) < 1e-10 and rel(out["Phi"][1], ref["Phi"]) < 1e-10
    hnd.close()
    with pytest.raises(r.RBMError):
        r.PICHandle(n, r.PoissonSolverPBSplines(4, nh, L), ctx)


@pytest.mark.parametrize("n,nh,nsamp", [(200_003, 64, 2), (700_001, 16, 1), (40_000, 30, 3)])
def test_chunk_sum_and_tile_assignment_switches(ctx, monkeypatch, n, nh, nsamp):
    """The step kernel's default path (chunk sums as 64-bit fixed-point integer atomics, tiles dealt to the CTAs
    round-robin, latency-organised field solve) against its alternatives (RBM_ACC_FIXED=0: per-chunk partial vectors summed in
    a fixed order; RBM_INTERLEAVE=0: contiguous chunks) and against the oracle: every combination is bit-reproducible from
    run to run, the combinations agree to round-off, and each meets the parity bars.  Sizes: several tiles per CTA with a
    ragged last tile and two samples in lock-step; the small grid's 512-thread kernel; a grid that is not a power of two."""
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=n % 1000)
    chis = np.linspace(0.7, 1.3, nsamp)
    nt, ns = 12, 4
    ref = [po.integrate(x, v, float(w[0]), L, nh, float(c), 0.1, nt, ns, extended=True, nthreads=po.max_threads()) for c in chis]
    outs = {}
    for acc in ("1", "0"):
        for il in ("1", "0"):
            monkeypatch.setenv("RBM_ACC_FIXED", acc)
            monkeypatch.setenv("RBM_INTERLEAVE", il)
            hnd = handle(ctx, n, nh)
            hnd.set_particles(x, v, float(w[0]))
            o = _run(hnd, chis, 0.1, nt, ns, n, nh)
            again = _run(hnd, chis, 0.1, nt, ns, n, nh)
            third = _run(hnd, chis, 0.1, nt, ns, n, nh)       # (from the second identical call on: CUDA-graph replay)
            hnd.close()
            for k in o:
                assert np.array_equal(o[k], again[k]) and np.array_equal(o[k], third[k]), (acc, il, k)
            for p in range(nsamp):
                assert rel(o["W"][p], ref[p]["W"]) < TOL_TRACE and rel(o["K"][p], ref[p]["K"]) < TOL_TRACE
                assert rel(o["Phi"][p], ref[p]["Phi"]) < 1e-10 and np.max(np.abs(o["X"][p] - ref[p]["X"])) < 1e-11
            outs[(acc, il)] = o
    base = outs[("1", "1")]
    for key, o in outs.items():
        for k in ("X", "V", "Phi", "W", "K", "M"):
            assert rel(o[k], base[k]) < 1e-12, (key, k)
