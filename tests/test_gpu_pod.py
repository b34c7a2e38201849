"""GPU parity tests of the snapshot POD (Gram + EVD), basis projection and DEIM through the C ABI against
the NumPy oracle and the reference's known-answer test (test/deim_test.jl).  Bar: POD singular values
<= 1e-10 relative on the well-conditioned leading range (lambda_i/lambda_1 >= 1e-6; a Gram-based EVD cannot
resolve smaller ones in FP64 -- SURVEY.md 7.3)."""
import os

import numpy as np
import pytest

from conftest import rel
from oracle import pod_numpy as pod

pytestmark = pytest.mark.gpu

import rbm_b200 as r  # noqa: E402


def _s(x, mu):
    return (1 - x) * np.cos(3 * np.pi * mu * (x + 1)) * np.exp(-(1 + x) * mu)


@pytest.mark.parametrize("n,c1,c2", [(5000, 300, 0), (4097, 130, 77), (20_000, 251, 251), (333, 5, 0), (64, 200, 0)])
def test_gram_vs_numpy(ctx, n, c1, c2):
    rng = np.random.default_rng(n)
    A = np.asfortranarray(rng.standard_normal((n, c1)))
    mats = [A]
    if c2:
        mats.append(np.asfortranarray(rng.standard_normal((n, c2))))
    G = r.gram(*mats, ctx=ctx)
    ref = pod.gram(*mats)
    assert rel(G, ref) < 1e-13
    assert np.array_equal(G, G.T)                                    # exactly symmetric (mirrored lower triangle)


def test_gram_strided_blocks_and_device_input(ctx):
    import torch

    rng = np.random.default_rng(0)
    big = np.asfortranarray(rng.standard_normal((1000, 64)))
    view = big[:900, :]                                              # ld = 1000 > nrows = 900
    assert rel(r.gram(view, ctx=ctx), view.T @ view) < 1e-13
    t = torch.from_numpy(np.ascontiguousarray(big.T)).cuda().T       # column-major device matrix
    assert t.stride(0) == 1
    assert rel(r.gram(t, ctx=ctx), big.T @ big) < 1e-13


def test_reference_deim_kat_on_gpu(ctx, golden_dir):
    """test/deim_test.jl:1-36 through the GPU path."""
    g = np.load(os.path.join(golden_dir, "pod_deim_kat.npz"))
    x = np.linspace(-1, 1, 100)
    mu_v = np.linspace(1, np.pi, 101)
    Psi, Lam = r.get_PODBasis_EVD(g["S"], k=20, ctx=ctx)
    assert Psi.shape == (100, 20) and Lam.shape == (51,)
    assert Lam[9] < 1e0 and Lam[14] < 1e-2 and Lam[19] < 1e-7       # deim_test.jl:14-18
    lead = g["Lam"] / g["Lam"][0] > 1e-6
    assert np.max(np.abs(Lam[lead] - g["Lam"][lead]) / g["Lam"][lead]) < 1e-10
    # basis vectors agree with the oracle's up to sign on the well-separated leading modes
    for j in range(8):
        d = min(np.linalg.norm(Psi[:, j] - g["Psi"][:, j]), np.linalg.norm(Psi[:, j] + g["Psi"][:, j]))
        assert d < 1e-7, j
    Pi = r.get_DEIM_interpolation_matrix(Psi, ctx=ctx)
    assert Pi.shape == (100, 20) and np.all(Pi.sum(axis=0) == 1.0) and len(set(np.argmax(Pi, axis=0))) == 20
    Sv = _s(x[:, None], mu_v[None, :])
    D = Psi @ np.linalg.inv(Pi.T @ Psi)
    xe = Pi.T @ x
    Sa = np.stack([D @ _s(xe, m) for m in mu_v], axis=1)
    assert np.mean(np.linalg.norm(Sv - Sa, axis=0)) < 5e-5          # deim_test.jl:34-36
    # same Psi in -> same points out as the literal restatement of deim.jl
    assert np.array_equal(r.deim_indices(g["Psi"], ctx=ctx), g["idx"])
    assert np.array_equal(r.deim_indices(Psi, ctx=ctx), pod.deim_indices(Psi))


def test_pod_singular_values_and_rank(ctx):
    rng = np.random.default_rng(3)
    n, m, rk = 30_000, 120, 40
    U, _ = np.linalg.qr(rng.standard_normal((n, rk)))
    Wm, _ = np.linalg.qr(rng.standard_normal((2 * m, rk)))
    sig = 10.0 ** (-np.arange(rk) / 6.0)
    S = np.asfortranarray(U @ np.diag(sig) @ Wm.T)
    X, V = np.asfortranarray(S[:, :m]), np.asfortranarray(S[:, m:])
    Psi, Lam = r.get_PODBasis_cotangentLiftEVD(X, V, tolerance=1e-8, ctx=ctx)
    Psi0, Lam0 = pod.pod_cotangent_lift_evd(X, V, tolerance=1e-8)
    assert Lam.shape == (2 * m,) and Psi.shape == Psi0.shape
    lead = sig ** 2 / sig[0] ** 2 > 1e-6
    sv = np.sqrt(np.abs(Lam[:rk]))
    assert np.max(np.abs(sv[lead] - sig[lead]) / sig[lead]) < 1e-10
    assert np.max(np.abs(Lam[:rk][lead] - Lam0[:rk][lead]) / Lam0[:rk][lead]) < 1e-10
    k = Psi.shape[1]
    assert rel(Psi.T @ Psi, np.eye(k)) < 1e-6
    assert rel(Psi @ (Psi.T @ S), Psi0 @ (Psi0.T @ S)) < 1e-7      # same subspace
    Psi5, _ = r.get_PODBasis_EVD(S, k=5, ctx=ctx)
    assert Psi5.shape == (n, 5)
    with pytest.raises(AssertionError):
        r.get_PODBasis_cotangentLiftEVD(X, V[:, :-1], ctx=ctx)
    with pytest.raises(r.RBMError):
        r.get_PODBasis_EVD(S, k=10_000, ctx=ctx)


def test_project_and_deim_larger(ctx):
    rng = np.random.default_rng(5)
    n, k, m = 50_000, 37, 301
    Q, _ = np.linalg.qr(rng.standard_normal((n, k)))
    Q = np.asfortranarray(Q)
    S = np.asfortranarray(rng.standard_normal((n, m)))
    assert rel(r.project(Q, S, ctx=ctx), Q.T @ S) < 1e-12
    assert np.array_equal(r.deim_indices(Q, ctx=ctx), pod.deim_indices(Q))


def test_reduced_basis_from_trainingset(ctx):
    """ReducedBasis(CotangentLiftEVD(), ts; particle_tol, field_tol)  (evd.jl:119-137) on a PIC training set."""
    L = r.BumpOnTail.domain_length(0.3)
    n, nh, nt = 4000, 16, 30
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=2)
    particles = r.ParticleList(x[None], v[None], w[None])
    poisson = r.PoissonSolverPBSplines(3, nh, L)
    pspace = r.ParameterSpace(r.Parameter("chi", 1 / 3, 5 / 3, 3))
    ip = r.IntegratorParameters(0.1, nt, nt + 1, nh, n, 3)
    ts = r.TrainingSet.create(particles, poisson, nt + 1, {"kappa": 0.3}, pspace, ip)
    r.integrate_vp_all(particles, poisson, pspace.column("chi"), ip, ts.snapshots, ctx=ctx)
    rb = r.ReducedBasis.from_trainingset(r.CotangentLiftEVD(), ts, particle_tol=1e-4, field_tol=1e-2, ctx=ctx)
    ref = pod.reduced_basis(ts.snapshots.matrix("X"), ts.snapshots.matrix("V"), ts.snapshots.matrix("A"), 1e-4, 1e-2)
    assert rb.k_p == ref["Psi_p"].shape[1] and rb.k_e == ref["Psi_e"].shape[1]
    lead = ref["Lam_p"] / ref["Lam_p"][0] > 1e-6
    assert np.max(np.abs(rb.Lam_p[lead] - ref["Lam_p"][lead]) / ref["Lam_p"][lead]) < 1e-10
    lead = ref["Lam_e"] / ref["Lam_e"][0] > 1e-6
    assert np.max(np.abs(rb.Lam_e[lead] - ref["Lam_e"][lead]) / ref["Lam_e"][lead]) < 1e-10
    XV = np.hstack([ts.snapshots.matrix("X"), ts.snapshots.matrix("V")])
    assert rel(rb.Psi_p @ (rb.Psi_p.T @ XV), ref["Psi_p"] @ (ref["Psi_p"].T @ XV)) < 1e-8
    assert rb.Pi_e.shape == (n, rb.k_e) and np.all(rb.Pi_e.sum(axis=0) == 1.0)
    # DEIM quality as the reference grades it: interpolation error of the field snapshots
    E = ts.snapshots.matrix("A")
    idx = np.argmax(rb.Pi_e, axis=0)
    D = rb.Psi_e @ np.linalg.inv(rb.Psi_e[idx, :])
    err = np.linalg.norm(E - D @ E[idx, :]) / np.linalg.norm(E)
    idx0 = ref["idx_e"]
    D0 = ref["Psi_e"] @ np.linalg.inv(ref["Psi_e"][idx0, :])
    err0 = np.linalg.norm(E - D0 @ E[idx0, :]) / np.linalg.norm(E)
    assert err < 3 * err0 + 1e-12


def test_tall_gram_known_spectrum(ctx):
    """Size-independent property at a tall shape: A = Q diag(sigma) W' built on the device; the Gram spectrum
    must return sigma^2 (no CPU SVD needed)."""
    import torch

    torch.manual_seed(1234)
    n, c = 1_000_000, 256
    Q, _ = torch.linalg.qr(torch.randn(n, c, dtype=torch.float64, device="cuda"))
    Wm, _ = torch.linalg.qr(torch.randn(c, c, dtype=torch.float64, device="cuda"))
    sig = 10.0 ** (-4.0 * torch.arange(c, dtype=torch.float64, device="cuda") / c)
    A = ((Q * sig) @ Wm.T).T.contiguous().T                          # column-major (n, c)
    assert A.stride(0) == 1
    Psi, Lam = r.get_PODBasis_EVD(A, k=8, ctx=ctx)
    s = sig.cpu().numpy()
    lead = s ** 2 / s[0] ** 2 > 1e-6
    assert np.max(np.abs(np.sqrt(Lam[lead]) - s[lead]) / s[lead]) < 1e-10
    assert Psi.is_cuda                                               # device matrix in -> device basis out
    assert rel((Psi.T @ Psi).cpu().numpy(), np.eye(8)) < 1e-8


def test_pod_edge_shapes(ctx):
    rng = np.random.default_rng(9)
    one = np.asfortranarray(rng.standard_normal((1000, 1)))                          # a single snapshot column
    Psi, Lam = r.get_PODBasis_EVD(one, k=1, ctx=ctx)
    assert Lam.shape == (1,) and abs(Lam[0] - float((one.T @ one)[0, 0])) < 1e-12 * Lam[0]
    assert rel(np.abs(Psi[:, 0]), np.abs(one[:, 0]) / np.linalg.norm(one)) < 1e-12
    S = np.asfortranarray(rng.standard_normal((257, 9)))                             # k == C, odd sizes (8-byte copies)
    Psi, Lam = r.get_PODBasis_EVD(S, k=9, ctx=ctx)
    assert rel(Psi.T @ Psi, np.eye(9)) < 1e-10
    assert np.array_equal(r.deim_indices(Psi, ctx=ctx), pod.deim_indices(Psi))
    wide = np.asfortranarray(rng.standard_normal((40, 300)))                         # more columns than rows
    G = r.gram(wide, ctx=ctx)
    assert rel(G, wide.T @ wide) < 1e-13
    with pytest.raises(r.RBMError):
        r.deim_indices(np.asfortranarray(rng.standard_normal((5, 9))), ctx=ctx)      # k > N
    with pytest.raises(ValueError):
        r.gram(S, wide, ctx=ctx)                                                     # DimensionMismatch


def test_device_resident_handover_pic_to_pod(ctx):
    """stage 1 -> stage 2 without a host round trip.  rbm_pic_run writes the snapshots into device
    arrays; the POD reads them where they lie (column-major (np, ns*nparam) views of the same memory) and agrees with
    the host path."""
    import torch

    L = r.BumpOnTail.domain_length(0.3)
    n, nh, nt, nparam = 6000, 16, 24, 4
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=3)
    hnd = r.PICHandle(n, r.PoissonSolverPBSplines(3, nh, L), ctx)
    hnd.set_particles(x, v, float(w[0]))
    chis = np.linspace(0.5, 1.5, nparam)
    dev = {k: torch.zeros((nparam, nt + 1, n), dtype=torch.float64, device="cuda") for k in "XVA"}
    hnd.run(chis, 0.1, nt, nt + 1, X=dev["X"], V=dev["V"], A=dev["A"])
    cols = nparam * (nt + 1)
    Xd, Vd, Ad = (dev[k].view(cols, n).T for k in "XVA")              # (np, ns*nparam), stride (1, np): Julia's reshape
    assert Xd.stride(0) == 1 and Xd.data_ptr() == dev["X"].data_ptr()
    Psi_p, Lam_p = r.get_PODBasis_cotangentLiftEVD(Xd, Vd, tolerance=1e-8, ctx=ctx)
    Psi_e, Lam_e = r.get_PODBasis_EVD(Ad, tolerance=1e-6, ctx=ctx)
    Xh, Vh, Ah = (dev[k].cpu().numpy().reshape(cols, n).T for k in "XVA")
    Psi_p0, Lam_p0 = pod.pod_cotangent_lift_evd(Xh, Vh, tolerance=1e-8)
    Psi_e0, Lam_e0 = pod.pod_evd(Ah, tolerance=1e-6)
    assert Psi_p.shape == Psi_p0.shape and Psi_e.shape == Psi_e0.shape
    for got, ref in ((Lam_p, Lam_p0), (Lam_e, Lam_e0)):
        lead = ref / ref[0] > 1e-6
        assert np.max(np.abs(got[lead] - ref[lead]) / ref[lead]) < 1e-10
    hnd.close()


def test_device_resident_handover(ctx):
    """stage 1 writes the snapshots into HBM, stage 2 (POD, DEIM, projection, reduced basis) reads
    them there; results are identical to the host round trip (same kernels, same data)."""
    import torch
    r.set_default_context(ctx)
    n, nh, nt, nparam = 2000, 16, 20, 3
    L = r.BumpOnTail.domain_length(0.3)
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=11)
    particles = r.ParticleList(x[None], v[None], w[None]); poisson = r.PoissonSolverPBSplines(3, nh, L)
    pspace = r.ParameterSpace(r.Parameter("chi", 0.8, 1.2, nparam))
    ip = r.IntegratorParameters(0.1, nt, nt + 1, nh, n, nparam)
    ts_h = r.TrainingSet.create(particles, poisson, nt + 1, {"kappa": 0.3}, pspace, ip)
    ts_d = r.TrainingSet.create(particles, poisson, nt + 1, {"kappa": 0.3}, pspace, ip, device="cuda")
    r.integrate_vp_all(particles, poisson, pspace.column("chi"), ip, ts_h.snapshots)
    r.integrate_vp_all(particles, poisson, pspace.column("chi"), ip, ts_d.snapshots)
    assert ts_d.snapshots.on_device and not ts_h.snapshots.on_device
    assert ts_d.snapshots == ts_h.snapshots                                  # bitwise, through to_host()
    rb_h = r.ReducedBasis.from_trainingset(r.CotangentLiftEVD(), ts_h)
    rb_d = r.ReducedBasis.from_trainingset(r.CotangentLiftEVD(), ts_d)
    assert isinstance(rb_d.Psi_p, torch.Tensor) and rb_d.Psi_p.is_cuda and rb_d.Pi_e.ndim == 1
    assert rb_d == rb_h
    Zh = r.project(rb_h.Psi_p, ts_h.snapshots.matrix("X"))
    Zd = r.project(rb_d.Psi_p, ts_d.snapshots.matrix("X"))
    assert Zd.is_cuda and np.array_equal(Zd.cpu().numpy(), Zh)
    pinned = ts_d.snapshots.to_host(pinned=True, non_blocking=True)
    torch.cuda.synchronize()
    assert pinned == ts_h.snapshots


@pytest.mark.parametrize("n", [400_000, 400_001])
def test_deim_many_rows(ctx, n):
    """Row counts that take the slab-per-warp residual kernel (>= 148*32 slabs of 64 rows), with 16-byte loads (even n)
    and scalar loads (odd n): same points as the literal restatement of deim.jl."""
    rng = np.random.default_rng(n)
    Q, _ = np.linalg.qr(rng.standard_normal((n, 12)))
    Q = np.asfortranarray(Q)
    idx = r.deim_indices(Q, ctx=ctx)
    assert np.array_equal(idx, pod.deim_indices(Q))
    assert len(set(idx.tolist())) == 12


def test_explicit_decomposition_handle_and_no_hidden_cache(ctx):
    """rbm_pod_factor / rbm_pod_basis: the decomposition lives in a caller-owned handle.  rbm_pod_evd keeps nothing between
    calls: a rank query followed by an in-place change of the data gives the basis of the NEW data (the
    pointer-keyed cache returned the stale one)."""
    import ctypes as C
    from rbm_b200 import _lib
    from rbm_b200.reducedbasis import _blocks

    rng = np.random.default_rng(3)
    S = np.asfortranarray(rng.standard_normal((3000, 40)) @ np.diag(0.5 ** np.arange(40)))
    Psi, Lam = r.get_PODBasis_EVD(S, tolerance=1e-6, ctx=ctx)           # factor -> basis through the handle
    Pref, Lref = pod.pod_evd(S, tolerance=1e-6)
    assert Psi.shape == Pref.shape
    lead = Lref / Lref[0] > 1e-6
    assert np.max(np.abs(Lam[lead] - Lref[lead]) / Lref[lead]) < 1e-10
    for j in range(min(5, Psi.shape[1])):
        assert min(np.linalg.norm(Psi[:, j] - Pref[:, j]), np.linalg.norm(Psi[:, j] + Pref[:, j])) < 1e-8
    # rank query, then overwrite the SAME buffer, then ask for the basis with the old entry point
    keep, P, N, Ld, nb, nrows, Ctot = _blocks((S,))
    lam = np.empty(Ctot); kout = C.c_int(0)
    ctx.check(ctx.lib.rbm_pod_evd(ctx.h, P, N, Ld, nb, nrows, 0, 1e-6, lam.ctypes.data, C.byref(kout), None, 0, 0))
    S[...] = np.asfortranarray(rng.standard_normal((3000, 40)) @ np.diag(0.7 ** np.arange(40)))
    k = 6
    out = np.empty((3000, k), order="F")
    ctx.check(ctx.lib.rbm_pod_evd(ctx.h, P, N, Ld, nb, nrows, k, 1e-6, lam.ctypes.data, C.byref(kout), _lib.ptr(out), 3000, k))
    Pnew, Lnew = pod.pod_evd(S, k=k)
    assert np.max(np.abs(lam[:k] - Lnew[:k]) / Lnew[:k]) < 1e-10
    for j in range(k):
        assert min(np.linalg.norm(out[:, j] - Pnew[:, j]), np.linalg.norm(out[:, j] + Pnew[:, j])) < 1e-8


def test_degenerate_snapshots_are_errors_not_faults(ctx):
    """All-zero snapshots (e.g. the A block of a reduced run, which save_solution never writes): the reference divides
    0/0; here rbm_pod_evd reports RBM_ERR_NUMERIC.  A NaN column handed to rbm_deim is reported too instead of
    faulting the CUDA context."""
    Z = np.zeros((500, 12), order="F")
    with pytest.raises(r.RBMError) as e:
        r.get_PODBasis_EVD(Z, ctx=ctx)
    assert e.value.code == -5
    rng = np.random.default_rng(1)
    Psi = np.asfortranarray(np.linalg.qr(rng.standard_normal((400, 6)))[0])
    Psi[:, 2] = np.nan
    with pytest.raises(r.RBMError) as e:
        r.deim_indices(Psi, ctx=ctx)
    assert e.value.code == -5
    good = np.asfortranarray(np.linalg.qr(rng.standard_normal((400, 6)))[0])      # the context is still usable
    assert np.array_equal(r.deim_indices(good, ctx=ctx), pod.deim_indices(good))


@pytest.mark.parametrize("n,k", [(3000, 150), (3001, 150), (96, 64), (4097, 33), (20_000, 97)])
def test_deim_blocked_panels(ctx, n, k):
    """The single-rank selection is one cooperative kernel, blocked in panels of 32 columns with a trailing update after
    each panel (pod.cu: k_deim_blocked).  Several panels, a ragged last panel, more than one 64-column tile of trailing
    columns, odd row counts (scalar path) and few rows: same points as the literal restatement of deim.jl:6-21, and the
    same points as the per-point launch path (RBM_DEIM_BLOCKED=0)."""
    rng = np.random.default_rng(n + k)
    Q, _ = np.linalg.qr(rng.standard_normal((n, k)))
    Q = np.asfortranarray(Q)
    idx = r.deim_indices(Q, ctx=ctx)
    assert np.array_equal(idx, pod.deim_indices(Q))
    os.environ["RBM_DEIM_BLOCKED"] = "0"
    try:
        assert np.array_equal(r.deim_indices(Q, ctx=ctx), idx)
    finally:
        del os.environ["RBM_DEIM_BLOCKED"]
    for _ in range(3):                       # no timing-dependent result (grid barriers, row ownership)
        assert np.array_equal(r.deim_indices(Q, ctx=ctx), idx)
