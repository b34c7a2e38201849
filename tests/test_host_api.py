"""CPU: host-side mirror of the reference API (mirrors test/parameter_tests.jl, parameterspace_tests.jl,
trainingset_tests.jl, reducedbasis_tests.jl) and the C-ABI contract (library loads, exports every symbol
declared in include/rbm_b200.h, fails loudly without a GPU)."""
import os
import re

import numpy as np
import pytest

import rbm_b200 as r
from rbm_b200.parameter import Parameter, ParameterSpace, CartesianParameterSampler, sample

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- test/parameter_tests.jl
def test_parameter_ctor_and_accessors():
    with pytest.raises(AssertionError):
        Parameter("p", 1.0, 0.0)                                   # minimum <= maximum (parameter.jl:21)
    with pytest.raises(AssertionError):
        Parameter("p", 0.0, 1.0, [0.5, 1.5])
    p = Parameter("p", 0.0, 1.0, [0.75, 0.25, 0.25, 0.5])
    assert list(p.samples) == [0.25, 0.5, 0.75]                    # sorted unique
    assert len(p) == 3 and p.size == (3,) and p[0] == 0.25 and p.hassamples()
    q = Parameter("q", 0.0, 1.0)
    assert len(q) == 0 and not q.hassamples()
    with pytest.raises(IndexError):
        q[0]
    assert Parameter("p", 0.0, 1.0, 5) == Parameter("p", 0.0, 1.0, np.linspace(0, 1, 5))
    assert Parameter("s", [0.2, 0.1]).minimum == 0.1
    assert hash(Parameter("p", 0.0, 1.0, 3)) == hash(Parameter("p", 0.0, 1.0, 3))
    assert Parameter("p", 0.0, 1.0, 3) != Parameter("p", 0.0, 1.0, 4)


# ---- test/parameterspace_tests.jl:2-43 (first parameter fastest)
def test_parameterspace_cartesian_order():
    p1 = Parameter("p1", 0.0, 1.0, [0.0, 0.5, 1.0])
    p2 = Parameter("p2", 0.0, 1.0, [0.5])
    p3 = Parameter("p3", 0.0, 2.0, [1.0, 2.0])
    ps = ParameterSpace(p1, p2, p3)
    grid = np.array([[0.0, 0.5, 1.0], [0.5, 0.5, 1.0], [1.0, 0.5, 1.0],
                     [0.0, 0.5, 2.0], [0.5, 0.5, 2.0], [1.0, 0.5, 2.0]])
    got = np.stack([ps.samples[k] for k in ("p1", "p2", "p3")], axis=1)
    assert np.array_equal(got, grid)
    assert ps == ParameterSpace(CartesianParameterSampler(), p1, p2, p3) == ParameterSpace(r.named_parameters(p1, p2, p3))
    assert len(ps) == 6 and ps.ndims() == 3 and ps.size == (6, 3)
    assert ps(4) == {"p1": 0.5, "p2": 0.5, "p3": 2.0}
    assert ParameterSpace.from_dict(ps.to_dict()) == ps
    assert list(sample(CartesianParameterSampler(), p1, p3)["p3"]) == [1.0, 1.0, 1.0, 2.0, 2.0, 2.0]


def test_training_script_parameter_space():
    """scripts/bump_on_tail_1_training.jl:43-48,61: chi in LinRange(1/3, 5/3, 10), other parameters fixed."""
    kappa = 0.3
    ps = ParameterSpace(Parameter("chi", 0.1 / kappa, 0.5 / kappa, 10), Parameter("eps", 0.03, 0.03, 1),
                        Parameter("a", 0.1, 0.1, 1), Parameter("v0", 4.5, 4.5, 1), Parameter("sigma", 0.5, 0.5, 1))
    assert len(ps) == 10
    assert np.allclose(ps.column("chi"), np.linspace(1 / 3, 5 / 3, 10))
    assert ps(9)["chi"] == pytest.approx(5 / 3)


# ---- test/trainingset_tests.jl:4-69
def test_trainingset_containers_and_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    particles = r.ParticleList(rng.random((1, 100)), rng.random((1, 100)), rng.random((1, 100)))
    poisson = r.PoissonSolverPBSplines(3, 10, 2 * np.pi)
    ps = ParameterSpace(Parameter("a", 0.0, 1.0, 3), Parameter("b", 0.0, 1.0, 1), Parameter("c", 0.0, 1.0, 2))
    ip = r.IntegratorParameters(0.1, 5, 6, 10, 100, len(ps))
    ts = r.TrainingSet.create(particles, poisson, 6, {"kappa": 0.3}, ps, ip)
    ss = ts.snapshots
    assert ss.X.shape == (6, 6, 100, 1) and ss.Phi.shape == (6, 6, 10, 1) and ss.W.shape == (6, 6)
    for k in ss.FIELDS:
        getattr(ss, k)[...] = rng.random(getattr(ss, k).shape)
    assert ss.matrix("X").shape == (100, 36) and ss.matrix("X").flags.f_contiguous
    assert ss.matrix("X")[7, 6 * 2 + 3] == ss.X[2, 3, 7, 0]          # column = ts + ns*p
    assert ts.initconds is particles and len(ts.initconds) == 100
    path = str(tmp_path / "temp.npz")
    ts.save(path)
    ts2 = r.TrainingSet.load(path)
    assert ts2.parameters == ts.parameters and ts2.paramspace == ts.paramspace
    assert np.array_equal(ts2.initconds.list, ts.initconds.list)
    assert ts2.snapshots == ts.snapshots and ts2.integrator == ts.integrator and ts2.poisson == ts.poisson
    assert ts2 == ts


def test_integrator_parameters_time_axis():
    ip = r.IntegratorParameters(0.1, 250, 251, 16, 10_000, 10)
    assert ip.t[0] == 0.0 and ip.t[-1] == pytest.approx(25.0) and len(ip.t) == 251   # not scaled by chi
    assert r.IntegratorParameters.from_attrs(ip.attrs()) == ip


def test_particle_list_views():
    x = np.arange(6.0).reshape(2, 3)
    pl = r.ParticleList(x, x + 10, np.ones((1, 3)))
    assert pl.list.shape == (5, 3) and len(pl) == 3
    pl.list[0, 0] = 42.0
    assert pl.x[0, 0] == 42.0                                         # views of one stacked array
    assert pl.uniform_weight() == 1.0
    with pytest.raises(ValueError):
        r.ParticleList(np.zeros((1, 3)), np.zeros((1, 4)), np.ones((1, 3)))


# ---- test/reducedbasis_tests.jl
def test_reducedbasis_container():
    rng = np.random.default_rng(1)
    rb = r.ReducedBasis(r.CotangentLiftEVD(), {"kappa": 0.3}, None, None, None, None,
                        rng.random(8), rng.random((50, 3)), rng.random(4), rng.random((50, 2)), rng.random((50, 2)))
    assert rb.k_p == 3 and rb.k_e == 2
    assert isinstance(r.CotangentLiftSVD(), r.ReductionAlgorithm)


def test_sorteigen_host():
    l, v = r.sorteigen(np.array([1.0, -3.0, 2.0]), np.eye(3))
    assert list(l) == [-3.0, 2.0, 1.0]
    with pytest.raises(TypeError):
        r.sorteigen(np.array([1.0 + 1j]), np.eye(1))


def test_sampler_properties():
    x, v, w = r.BumpOnTail.draw_accept_reject(20_000, seed=1)
    L = r.BumpOnTail.domain_length(0.3)
    assert x.min() >= 0 and x.max() < L and np.all(w == L / 20_000)
    assert abs(v.mean() - 0.45) < 0.05                              # a * v0
    assert abs(np.mean(np.cos(0.3 * x)) - 0.015) < 0.02              # eps/2 modulation
    x2, _, _ = r.BumpOnTail.draw_accept_reject(20_000, seed=1)
    assert np.array_equal(x, x2)


# ---- C-ABI contract
def test_library_exports_every_declared_symbol():
    lib = r._lib.load()
    header = open(os.path.join(ROOT, "include", "rbm_b200.h")).read()
    declared = set(re.findall(r"\b(rbm_[a-z0-9_]+)\s*\(", header))
    declared -= {"rbm_ctx", "rbm_pic"}
    assert declared, "no declarations parsed"
    assert declared == set(r.EXPORTS), declared ^ set(r.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in rbm_b200.h but not exported"
    assert lib.rbm_version() == 100


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, never route through the oracle."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(r.RBMError) as e:
        r.Context(0)
    assert "no CPU fallback" in str(e.value)
    import rbm_b200.poisson as pz, rbm_b200.integrator as iz, rbm_b200.reducedbasis as rz
    for mod in (pz, iz, rz, r._lib):
        src = open(mod.__file__).read()
        assert "oracle" not in src, f"{mod.__name__} must not reference the oracle"


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """The boundary is a C ABI: include/rbm_b200.h must compile as C99 (no C++ leaks) and a C program must link
    against the shared library and call an entry point that needs no GPU."""
    import shutil
    import subprocess

    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "rbm_b200.h"\n'
                   'int main(void) { printf("%d\\n", rbm_version()); return rbm_version() > 0 ? 0 : 1; }\n')
    exe = tmp_path / "abi"
    libdir = os.path.join(root, "reducedbasismethods.jl_b200", "lib")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"), str(src),
                    "-o", str(exe), "-L", libdir, "-lrbm_b200", f"-Wl,-rpath,{libdir}"], check=True, env=env)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True, env=env)
    assert int(out.stdout.strip()) > 0


def test_snapshots_host_container_and_reduced_basis_views():
    """Host-side behaviour of the containers that also have a device-resident mode (no GPU needed for the host mode)."""
    ip = r.IntegratorParameters(0.1, 10, 11, 8, 5, 3)
    ss = r.Snapshots.from_integrator(ip)
    assert not ss.on_device and ss.to_host() is ss
    assert ss.X.shape == (3, 11, 5, 1) and ss.Phi.shape == (3, 11, 8, 1) and ss.W.shape == (3, 11)
    ss.X[...] = np.arange(ss.X.size, dtype=np.float64).reshape(ss.X.shape)
    M = ss.matrix("X")                                     # reshape(SS.X, (np, ns*nparam)), column-major view
    assert M.shape == (5, 33) and M[2, 12] == ss.X[1, 1, 2, 0] and not M.flags.owndata
    assert set(ss.to_dict()) == set(r.Snapshots.FIELDS)
    assert ss == r.Snapshots(arrays=ss.to_dict())
    # ReducedBasis keeps either the dense 0/1 DEIM matrix (reference layout) or the selected rows; host() unifies them
    Psi = np.linalg.qr(np.random.default_rng(0).standard_normal((20, 3)))[0]
    idx = np.array([4, 0, 17])
    Pi = np.zeros((20, 3)); Pi[idx, np.arange(3)] = 1.0
    a = r.ReducedBasis(r.CotangentLiftEVD(), {}, None, None, None, None, np.ones(6), Psi, np.ones(3), Psi, Pi)
    b = r.ReducedBasis(r.CotangentLiftEVD(), {}, None, None, None, None, np.ones(6), Psi, np.ones(3), Psi, idx)
    assert a == b and np.array_equal(b.host("Pi_e"), Pi) and b.k_e == 3


def test_check_buffer_rules():
    """What the host mirror refuses to hand to the C ABI as an output (raw addresses are not bounds-checked)."""
    from rbm_b200._lib import check_buffer

    check_buffer(None, 10, "X")
    check_buffer(np.zeros((2, 5)), 10, "X")
    check_buffer(np.zeros((5, 2), order="F"), 10, "X")
    for bad in (np.zeros(10, dtype=np.float32), np.zeros(9), np.zeros(20)[::2]):
        with pytest.raises(ValueError):
            check_buffer(bad, 10, "X")
    ro = np.zeros(10); ro.flags.writeable = False
    with pytest.raises(ValueError):
        check_buffer(ro, 10, "X")
    check_buffer(ro, 10, "x", writable=False)
    with pytest.raises(TypeError):
        check_buffer([0.0] * 10, 10, "X")


def test_reducedbasis_save_load_roundtrip(tmp_path):
    """h5save(rb) / ReducedBasis(h5) (src/reducedbasis.jl:43-88): the whole tree comes back, so stage 3 can be driven from the
    projections file alone (scripts/bump_on_tail_3_testing.jl:25-42); mirrors test/reducedbasis_tests.jl."""
    rng = np.random.default_rng(0)
    n, nh = 50, 8
    pl = r.ParticleList(rng.random((1, n)), rng.random((1, n)), np.full((1, n), 1.0 / n))
    ps2 = ParameterSpace(Parameter("chi", 0.5, 1.5, 3))
    ip = r.IntegratorParameters(0.1, 10, 11, nh, n, 3)
    poisson = r.PoissonSolverPBSplines(3, nh, 2.0)
    Pi = np.zeros((n, 3)); Pi[[4, 9, 1], [0, 1, 2]] = 1.0
    rb = r.ReducedBasis(r.CotangentLiftEVD(), {"kappa": 0.3, "eps": 0.03}, ps2, pl, ip, poisson, rng.random(6),
                        rng.random((n, 4)), rng.random(5), rng.random((n, 3)), Pi)
    f = str(tmp_path / "rb.npz")
    rb.save(f)
    back = r.ReducedBasis.load(f)
    assert back == rb and back.k_p == 4 and back.k_e == 3
    assert back.parameters == rb.parameters and back.paramspace == ps2 and back.initconds == pl
    assert back.integrator == ip and back.poisson == poisson
    assert isinstance(back.algorithm, r.UnspecifiedAlgorithm)
    # a TrainingSet whose parameter space holds a Parameter without samples saves and loads (np.asarray(None) did not)
    from rbm_b200.parameter import OrderedDict
    psx = ParameterSpace(OrderedDict(chi=Parameter("chi", 0.5, 1.5, 3), free=Parameter("free", 0.0, 1.0)),
                         OrderedDict(chi=np.linspace(0.5, 1.5, 3), free=np.zeros(3)))
    ts = r.TrainingSet(dict(kappa=0.3), psx, pl, r.Snapshots(1, n, nh, 11, 3), ip, poisson)
    g = str(tmp_path / "ts.npz")
    ts.save(g)
    assert r.TrainingSet.load(g) == ts
