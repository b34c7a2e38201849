"""GPU: the reduced online model with DEIM (src/particles/electric_field.jl:2-43, time_marching.jl:108-150) through the
C ABI against the literal NumPy restatement (oracle/reduced_numpy.py), on a basis built by the GPU POD/DEIM path."""
import numpy as np
import pytest

from conftest import rel
from oracle import reduced_numpy as rn

pytestmark = pytest.mark.gpu

import rbm_b200 as r  # noqa: E402


@pytest.fixture(scope="module", params=[3000, 2501])     # odd N: padded leading dimensions, 8-byte edge copies
def setup(ctx, request):
    L = r.BumpOnTail.domain_length(0.3)
    n, nh, nt = request.param, 16, 40
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=5)
    particles = r.ParticleList(x[None], v[None], w[None])
    poisson = r.PoissonSolverPBSplines(3, nh, L)
    train = np.array([0.6, 1.0, 1.4])
    ip = r.IntegratorParameters(0.1, nt, nt + 1, nh, n, 3)
    ts = r.TrainingSet.create(particles, poisson, nt + 1, {"kappa": 0.3}, r.ParameterSpace(r.Parameter("chi", train)), ip)
    r.integrate_vp_all(particles, poisson, train, ip, ts.snapshots, ctx=ctx)
    rb = r.ReducedBasis.from_trainingset(r.CotangentLiftEVD(), ts, particle_tol=1e-10, field_tol=1e-8, ctx=ctx)
    return dict(L=L, n=n, nh=nh, nt=nt, x=x, v=v, w=w, particles=particles, poisson=poisson, rb=rb)


def test_reduced_model_matches_reference_restatement(ctx, setup):
    s = setup
    rb = s["rb"]
    idx = np.argmax(rb.Pi_e, axis=0)
    rfield = r.DEIMElectricField(s["poisson"], rb.Psi_p, rb.Psi_e, rb.Pi_e)
    assert np.array_equal(rfield.idx, idx)
    chis = np.array([0.8, 1.2])
    ip = r.IntegratorParameters(0.1, s["nt"], s["nt"] + 1, s["nh"], s["n"], 2)
    rss, Zx, Zv = r.reduced_integrate_vp_all(s["particles"], rfield, chis, ip, ctx=ctx, return_z=True)
    for p, chi in enumerate(chis):
        f = rn.deim_field(rn.ScaledField(s["L"], s["nh"], float(chi)), rb.Psi_p, rb.Psi_e, idx)
        o = rn.reduced_integrate(s["x"], s["v"], s["w"], rb.Psi_p, f, float(chi), 0.1, s["nt"], s["nt"] + 1)
        assert rel(Zx[p], o["Zx"]) < 1e-9 and rel(Zv[p], o["Zv"]) < 1e-8
        assert rel(rss.X[p, :, :, 0], o["X"]) < 1e-9 and rel(rss.V[p, :, :, 0], o["V"]) < 1e-8
        assert rel(rss.Phi[p, :, :, 0], o["Phi"]) < 1e-7
        assert rel(rss.W[p], o["W"]) < 1e-8 and rel(rss.K[p], o["K"]) < 1e-9 and rel(rss.M[p], o["M"]) < 1e-9
        assert np.all(rss.A[p] == 0.0)                          # save_solution never writes SS.A (time_marching.jl:86-101)
    # the reduced model tracks the full model it was trained around
    full = r.integrate_vp_all(s["particles"], s["poisson"], chis, ip, ctx=ctx)
    assert np.max(np.abs(rss.X - full.X)) < 0.2 and rel(rss.W, full.W) < 0.1
    # per-sample signature writes into slot p of the snapshots
    RSS = r.Snapshots.from_integrator(ip)
    r.reduced_integrate_vp(s["particles"], rb.Psi_p, {"chi": 1.2}, rfield, RSS, ip, 1, ctx=ctx)
    assert np.array_equal(RSS.X[1], rss.X[1]) and np.all(RSS.X[0] == 0.0)


def test_reduced_model_errors(ctx, setup):
    s = setup
    rb = s["rb"]
    with pytest.raises(ValueError):
        r.DEIMElectricField(s["poisson"], rb.Psi_p, rb.Psi_e[:-1], rb.Pi_e)
    with pytest.raises(ValueError):
        r.DEIMElectricField(s["poisson"], rb.Psi_p, rb.Psi_e, rb.Pi_e[:, :-1])
    bad = np.argmax(rb.Pi_e, axis=0).copy()
    bad[0] = s["n"] + 5
    with pytest.raises(r.RBMError):
        r.ReducedModel(s["particles"], r.DEIMElectricField(s["poisson"], rb.Psi_p, rb.Psi_e, bad), ctx)


def test_reduced_long_run_tracks_restatement(ctx):
    """250 steps into the nonlinear phase: the GPU reduced model stays on the trajectory of the NumPy restatement
    (round-off grows with the instability, so the bar is looser than for 40 steps); it is the reduced MODEL that
    leaves the full-order trajectory at late times, identically in both implementations."""
    L = r.BumpOnTail.domain_length(0.3)
    n, nh, nt = 3000, 16, 250
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=5)
    particles = r.ParticleList(x[None], v[None], w[None]); poisson = r.PoissonSolverPBSplines(3, nh, L)
    train = np.linspace(1 / 3, 5 / 3, 6)
    ip = r.IntegratorParameters(0.1, nt, 26, nh, n, len(train))
    ts = r.TrainingSet.create(particles, poisson, 26, {"kappa": 0.3}, r.ParameterSpace(r.Parameter("chi", train)), ip)
    r.integrate_vp_all(particles, poisson, train, ip, ts.snapshots, ctx=ctx)
    rb = r.ReducedBasis.from_trainingset(r.CotangentLiftEVD(), ts, particle_tol=1e-9, field_tol=1e-5, ctx=ctx)
    idx = np.argmax(rb.Pi_e, axis=0)
    rfield = r.DEIMElectricField(poisson, rb.Psi_p, rb.Psi_e, rb.Pi_e)
    chi = 1.1
    ip1 = r.IntegratorParameters(0.1, nt, 26, nh, n, 1)
    rss = r.reduced_integrate_vp_all(particles, rfield, [chi], ip1, ctx=ctx)
    f = rn.deim_field(rn.ScaledField(L, nh, chi), rb.Psi_p, rb.Psi_e, idx)
    o = rn.reduced_integrate(x, v, w, rb.Psi_p, f, chi, 0.1, nt, 26)
    assert rel(rss.W[0], o["W"]) < 1e-5 and rel(rss.K[0], o["K"]) < 1e-6
    assert rel(rss.X[0, :, :, 0], o["X"]) < 1e-5


def test_fused_reconstruct_deposit(ctx, monkeypatch):
    """on steps that do not save, the deposit runs as the epilogue of X = Psi Z (csrc/recon.cu) and X is never
    stored.  Shape chosen so that the fused kernel is taken (>= one tile per SM), with a ragged last row tile (odd N) and a
    second sample tile holding two samples.  Checked against the two-kernel path of the same library (summation order of
    the deposit differs: 1e-11) and against the NumPy restatement for two samples."""
    L = r.BumpOnTail.domain_length(0.3)
    n, nh, nt, ns = 20001, 16, 12, 4
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=11)
    particles = r.ParticleList(x[None], v[None], w[None]); poisson = r.PoissonSolverPBSplines(3, nh, L)
    train = np.array([0.6, 1.0, 1.4])
    ip = r.IntegratorParameters(0.1, nt, nt + 1, nh, n, 3)
    ts = r.TrainingSet.create(particles, poisson, nt + 1, {"kappa": 0.3}, r.ParameterSpace(r.Parameter("chi", train)), ip)
    r.integrate_vp_all(particles, poisson, train, ip, ts.snapshots, ctx=ctx)
    rb = r.ReducedBasis.from_trainingset(r.CotangentLiftEVD(), ts, particle_tol=1e-10, field_tol=1e-8, ctx=ctx)
    idx = np.argmax(rb.Pi_e, axis=0)
    rfield = r.DEIMElectricField(poisson, rb.Psi_p, rb.Psi_e, rb.Pi_e)
    chis = np.linspace(0.5, 1.5, 130)

    def run(fused, profile=False):
        monkeypatch.setenv("RBM_RM_FUSE_RECON", "1" if fused else "0")
        model = r.ReducedModel(particles, rfield, ctx)
        out = {k: np.zeros((len(chis), ns, rfield.kp)) for k in ("Zx", "Zv")}
        out["Phi"] = np.zeros((len(chis), ns, nh)); out["W"] = np.zeros((len(chis), ns))
        if profile:
            ctx.profile_enable(True)
        model.run(chis, 0.1, nt, ns, **out)
        launches = ctx.profile_query("k_recon_deposit")[0] if profile else None
        if profile:
            ctx.profile_enable(False)
        model.close()
        return out, launches

    fused, launches = run(True, profile=True)
    assert launches == nt                                   # every time step, not the saves (they need X)
    again, _ = run(True)
    for k in fused:
        assert np.array_equal(fused[k], again[k])           # fixed summation order: bit-reproducible
    plain, _ = run(False)
    for k in ("Zx", "Zv", "Phi", "W"):
        assert rel(fused[k], plain[k]) < 1e-11, k
    for p in (0, 129):                                      # first sample of tile 0, last sample of the ragged tile 1
        chi = float(chis[p])
        f = rn.deim_field(rn.ScaledField(L, nh, chi), rb.Psi_p, rb.Psi_e, idx)
        o = rn.reduced_integrate(x, v, w, rb.Psi_p, f, chi, 0.1, nt, ns)
        assert rel(fused["Zx"][p], o["Zx"]) < 1e-9 and rel(fused["Zv"][p], o["Zv"]) < 1e-8
        assert rel(fused["W"][p], o["W"]) < 1e-8 and rel(fused["Phi"][p], o["Phi"]) < 1e-7
