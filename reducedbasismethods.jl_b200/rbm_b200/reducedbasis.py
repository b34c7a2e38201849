"""ReducedBasis + POD/DEIM -- host API kept from the reference, compute in librbm_b200.

  ReductionAlgorithm tags .................. src/reducedbasis.jl:2-7
  ReducedBasis container ................... src/reducedbasis.jl:9-41
  get_PODBasis_cotangentLiftEVD(X, V) ...... src/algorithms/evd.jl:26-55
  get_PODBasis_EVD(S) ...................... src/algorithms/evd.jl:63-87
  get_DEIM_interpolation_matrix(Psi) ....... src/algorithms/deim.jl:6-21
  sorteigen ................................ src/eigen.jl:4-7
  ReducedBasis(CotangentLiftEVD(), ts) ..... src/algorithms/evd.jl:119-137

Matrices are passed in column-major memory: a NumPy array is taken as-is when it is
Fortran-ordered (e.g. ``Snapshots.matrix``) and copied once otherwise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .engine import default_context
from .trainingset import TrainingSet


class ReductionAlgorithm:
    pass


class UnspecifiedAlgorithm(ReductionAlgorithm):
    pass


class CotangentLiftEVD(ReductionAlgorithm):
    pass


class CotangentLiftSVD(ReductionAlgorithm):   # declared only in the reference too (no method)
    pass


def sorteigen(evals, evecs):
    """src/eigen.jl:4-7 (host helper; the library applies the same permutation on the device)."""
    evals = np.asarray(evals)
    if np.iscomplexobj(evals):
        raise TypeError("sorteigen accepts real eigenvalues only")
    p = np.argsort(-np.abs(evals), kind="stable")
    return evals[p], np.asarray(evecs)[:, p]


def _colmajor(a):
    """(array whose memory is column-major for the logical 2-D matrix, nrows, ncols, ld)."""
    if hasattr(a, "data_ptr"):          # torch tensor: logical (nrows, ncols) with stride (1, ld)
        if a.dim() != 2:
            raise ValueError("DimensionMismatch: expected a matrix")
        if not _lib.is_f64(a):
            raise ValueError(f"device matrices must be float64 (got {a.dtype})")
        nrows, ncols = a.shape
        if a.stride(0) != 1:
            raise ValueError("device matrices must be column-major (stride (1, ld))")
        return a, int(nrows), int(ncols), int(a.stride(1)) if ncols > 1 else int(nrows)
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 2:
        raise ValueError("DimensionMismatch: expected a matrix")
    if not a.flags.f_contiguous:
        a = np.asfortranarray(a)
    return a, a.shape[0], a.shape[1], a.shape[0]


def _blocks(mats):
    keep, ptrs, ncols, lds = [], [], [], []
    nrows = None
    for m in mats:
        a, r, c, ld = _colmajor(m)
        if nrows is None:
            nrows = r
        elif nrows != r:
            raise ValueError("DimensionMismatch: column blocks differ in row count")
        keep.append(a); ptrs.append(_lib.ptr(a)); ncols.append(c); lds.append(ld)
    nb = len(mats)
    P = (C.c_void_p * nb)(*ptrs)
    N = (C.c_int64 * nb)(*ncols)
    L = (C.c_int64 * nb)(*lds)
    return keep, P, N, L, nb, nrows, sum(ncols)


def gram(*mats, ctx=None):
    """G = [B1 B2 ...]'[B1 B2 ...] (C x C)."""
    ctx = ctx or default_context()
    keep, P, N, L, nb, nrows, Ctot = _blocks(mats)
    G = np.empty((Ctot, Ctot), dtype=np.float64, order="F")
    ctx.check(ctx.lib.rbm_gram(ctx.h, P, N, L, nb, nrows, G.ctypes.data))
    return G


def _on_device(a) -> bool:
    return hasattr(a, "data_ptr") and bool(getattr(a, "is_cuda", False))


def _empty_colmajor(nrows, ncols, like=None):
    """Column-major (nrows x ncols) float64 output: a torch CUDA tensor view when `like` lives on the device."""
    if like is not None and _on_device(like):
        import torch
        return torch.empty((ncols, nrows), dtype=torch.float64, device=like.device).T
    return np.empty((nrows, ncols), dtype=np.float64, order="F")


def _pod(mats, k, tolerance, ctx):
    """Device matrices in -> device basis out (eigenvalues always on the host: the rank decision is made there)."""
    ctx = ctx or default_context()
    keep, P, N, L, nb, nrows, Ctot = _blocks(mats)
    Lam = np.empty(Ctot, dtype=np.float64)
    kout = C.c_int(0)
    # factor (Gram + eigen + rank) -> allocate Psi with exactly that many columns -> basis; the decomposition lives in
    # an explicit handle between the two calls (nothing is cached inside the library)
    f = C.c_void_p()
    ctx.check(ctx.lib.rbm_pod_factor(ctx.h, P, N, L, nb, nrows, int(k), float(tolerance), Lam.ctypes.data,
                                     C.byref(kout), C.byref(f)))
    try:
        k = kout.value
        Psi = _empty_colmajor(nrows, k, like=keep[0])
        ctx.check(ctx.lib.rbm_pod_basis(f, P, N, L, nb, nrows, int(k), _lib.ptr(Psi), nrows))
    finally:
        ctx.lib.rbm_evd_destroy(f)
    return Psi, Lam


def get_PODBasis_cotangentLiftEVD(X, V, k=0, tolerance=1e-8, ctx=None):
    """-> (Psi (N x k), Lambda (all 2m eigenvalues))."""
    if tuple(X.shape) != tuple(V.shape):
        raise AssertionError("size(X) == size(V)")          # evd.jl:27
    return _pod((X, V), k, tolerance, ctx)


def get_PODBasis_EVD(S, k=0, tolerance=1e-4, ctx=None):
    return _pod((S,), k, tolerance, ctx)


def deim_indices(Psi, ctx=None):
    """Selected rows (0-based) of the greedy DEIM with the reference's signed arg-max."""
    ctx = ctx or default_context()
    a, n, k, ld = _colmajor(Psi)
    idx = np.empty(k, dtype=np.int64)
    ctx.check(ctx.lib.rbm_deim(ctx.h, _lib.ptr(a), n, ld, k, idx.ctypes.data))
    return idx


def get_DEIM_interpolation_matrix(Psi, ctx=None):
    """Dense N x k 0/1 matrix exactly as the reference returns it (deim.jl:9,18)."""
    idx = deim_indices(Psi, ctx)
    Pi = np.zeros((Psi.shape[0], Psi.shape[1]), dtype=np.float64, order="F")
    Pi[idx, np.arange(Psi.shape[1])] = 1.0
    return Pi


def project(Psi, S, ctx=None):
    """Z = Psi' S  (scripts/bump_on_tail_2_projections.jl:44-60)."""
    ctx = ctx or default_context()
    a, n, k, lda = _colmajor(Psi)
    b, n2, m, ldb = _colmajor(S)
    if n != n2:
        raise ValueError("DimensionMismatch")
    Z = _empty_colmajor(k, m, like=b if _on_device(b) else None)
    ctx.check(ctx.lib.rbm_project(ctx.h, _lib.ptr(a), n, lda, k, _lib.ptr(b), ldb, m, _lib.ptr(Z), k))
    return Z


class ReducedBasis:
    """Fields Lam_p, k_p, Psi_p, Lam_e, k_e, Psi_e, Pi_e (src/reducedbasis.jl:21-29)."""

    def __init__(self, algorithm, parameters, paramspace, initconds, integrator, poisson, Lam_p, Psi_p, Lam_e, Psi_e,
                 Pi_e):
        self.algorithm = algorithm
        self.parameters = parameters
        self.paramspace = paramspace
        self.initconds = initconds
        self.integrator = integrator
        self.poisson = poisson
        keep = lambda a: a if _on_device(a) else np.asarray(a)       # device bases stay in HBM
        self.Lam_p = np.asarray(Lam_p); self.Psi_p = keep(Psi_p); self.k_p = int(self.Psi_p.shape[1])
        self.Lam_e = np.asarray(Lam_e); self.Psi_e = keep(Psi_e); self.k_e = int(self.Psi_e.shape[1])
        self.Pi_e = np.asarray(Pi_e)     # dense N x k_e 0/1 matrix (reference layout) or the k_e selected rows

    @classmethod
    def from_trainingset(cls, alg: CotangentLiftEVD, ts: TrainingSet, particle_tol=1e-8, field_tol=1e-4, ctx=None):
        """ReducedBasis(alg::CotangentLiftEVD, ts; particle_tol, field_tol)  (evd.jl:119-137)."""
        if not isinstance(alg, CotangentLiftEVD):
            raise NotImplementedError("only CotangentLiftEVD has a method (as in the reference)")
        X = ts.snapshots.matrix("X"); V = ts.snapshots.matrix("V"); E = ts.snapshots.matrix("A")
        Psi_p, Lam_p = get_PODBasis_cotangentLiftEVD(X, V, tolerance=particle_tol, ctx=ctx)
        Psi_e, Lam_e = get_PODBasis_EVD(E, tolerance=field_tol, ctx=ctx)
        # device-resident training set: keep the bases in HBM and the DEIM selection as row indices (the dense
        # N x k_e 0/1 matrix of the reference is a host-side convenience; DEIMElectricField accepts either form)
        Pi_e = deim_indices(Psi_e, ctx=ctx) if ts.snapshots.on_device else get_DEIM_interpolation_matrix(Psi_e, ctx=ctx)
        return cls(alg, ts.parameters, ts.paramspace, ts.initconds, ts.integrator, ts.poisson, Lam_p, Psi_p, Lam_e,
                   Psi_e, Pi_e)

    def host(self, name):
        """Host (NumPy) view of a field, whatever memory it lives in; Pi_e always as the dense 0/1 matrix."""
        a = getattr(self, name)
        if _on_device(a):
            a = a.cpu().numpy()
        a = np.asarray(a)
        if name == "Pi_e" and a.ndim == 1:
            d = np.zeros((int(self.Psi_e.shape[0]), self.k_e), dtype=np.float64, order="F")
            d[a, np.arange(self.k_e)] = 1.0
            a = d
        return a

    def __eq__(self, o):
        return (isinstance(o, ReducedBasis) and self.k_p == o.k_p and self.k_e == o.k_e
                and all(np.array_equal(self.host(k), o.host(k)) for k in ("Lam_p", "Psi_p", "Lam_e", "Psi_e", "Pi_e")))

    # persistence: the tree of the reference's h5save / ReducedBasis(h5) (src/reducedbasis.jl:43-88): kp, ke, Λp, Ψp, Λe,
    # Ψe, Πe + groups parameters / parameterspace / initial_conditions / integrator + poisson attrs p, L.  No HDF5
    # toolchain in this image, so the same tree is a NumPy .npz with '/'-separated keys and ASCII dataset names (Lp = Λp,
    # Pp = Ψp, Le = Λe, Pe = Ψe, Pie = Πe).
    def save(self, path):
        d = dict(kp=np.asarray(self.k_p), ke=np.asarray(self.k_e), Lp=self.Lam_p, Pp=self.host("Psi_p"), Le=self.Lam_e,
                 Pe=self.host("Psi_e"), Pie=self.host("Pi_e"))
        d.update(_tree_of(self.parameters, self.paramspace, self.initconds, self.integrator, self.poisson))
        np.savez(path, **d)

    @classmethod
    def load(cls, path, algorithm=None):
        """ReducedBasis(fpath)  (src/reducedbasis.jl:43-66): the algorithm tag is not stored by the reference either
        (it reads back as UnspecifiedAlgorithm)."""
        z = np.load(path)
        params, pspace, particles, ip, poisson = _tree_from(z)
        return cls(algorithm or UnspecifiedAlgorithm(), params, pspace, particles, ip, poisson, z["Lp"], z["Pp"],
                   z["Le"], z["Pe"], z["Pie"])


def _tree_of(parameters, paramspace, initconds, integrator, poisson):
    """parameters / parameterspace / initial_conditions / integrator groups + p, L (shared with TrainingSet.save)."""
    d = {}
    for k, v in (parameters or {}).items():
        d[f"parameters/{k}"] = np.asarray(v)
    ps = paramspace.to_dict()
    for k, v in ps["samples"].items():
        d[f"parameterspace/samples/{k}"] = v
    for k, v in ps["parameters"].items():
        d[f"parameterspace/parameters/{k}/minimum"] = np.asarray(v["minimum"])
        d[f"parameterspace/parameters/{k}/maximum"] = np.asarray(v["maximum"])
        if v["samples"] is not None:               # a Parameter without samples stores no dataset (parameter.jl:96-104)
            d[f"parameterspace/parameters/{k}/samples"] = np.asarray(v["samples"])
    d["initial_conditions/list"] = initconds.list
    for k, v in integrator.attrs().items():
        d[f"integrator/{k}"] = np.asarray(v)
    d["p"] = np.asarray(poisson.p)
    d["L"] = np.asarray(poisson.L)
    return d


def _tree_from(z):
    from .parameter import ParameterSpace
    from .particles import ParticleList
    from .poisson import PoissonSolverPBSplines
    from .snapshots import IntegratorParameters

    params = {k.split("/", 1)[1]: float(z[k]) for k in z.files if k.startswith("parameters/")}
    names = [k.rsplit("/", 1)[1] for k in z.files if k.startswith("parameterspace/samples/")]
    psd = {"samples": {n: z[f"parameterspace/samples/{n}"] for n in names},
           "parameters": {n: {"minimum": float(z[f"parameterspace/parameters/{n}/minimum"]),
                              "maximum": float(z[f"parameterspace/parameters/{n}/maximum"]),
                              "samples": (z[f"parameterspace/parameters/{n}/samples"]
                                          if f"parameterspace/parameters/{n}/samples" in z.files else None)}
                          for n in names}}
    pl = z["initial_conditions/list"]
    nd = (pl.shape[0] - 1) // 2
    particles = ParticleList(pl[:nd], pl[nd:2 * nd], pl[2 * nd:])
    ip = IntegratorParameters.from_attrs({k: z[f"integrator/{k}"].item() for k in ("dt", "nt", "ns", "nh", "np", "nparam")})
    poisson = PoissonSolverPBSplines(int(z["p"]), ip.nh, float(z["L"]))   # nh from integrator attrs (poisson.jl:11-21)
    return params, ParameterSpace.from_dict(psd), particles, ip, poisson
