#=
tools/reference_vectors.jl -- turns "PIC parity unpinned" into a pinned oracle.

The PIC arithmetic of ReducedBasisMethods.jl lives in VlasovMethods.jl / PoissonSolvers.jl / ParticleMethods.jl
(Project.toml:19,21,26,36,38,41), which are not in /root/reference and cannot run in the build image (no Julia).
A maintainer with a Julia environment in which `using ReducedBasisMethods` works runs, from the repository root,

    julia --project=<env with ReducedBasisMethods> tools/reference_vectors.jl

It reads tests/golden/ref_julia_inputs.bin (the inputs of tests/golden/pic_small.npz: 2000 particles, nh = 16,
nt = 40, ns = 11, three chi; written by tools/export_reference_inputs.py), runs the REAL reference path

    PoissonSolverPBSplines(p, nh, L)                   scripts/bump_on_tail_1_training.jl:73
    ScaledPoissonField(poisson, chi)                   :94
    update!(efield, x, w, t); efield!(efield, E, x)    src/particles/electric_field.jl:22-31 (protocol)
    coefficients(efield); energy(efield)               src/particles/time_marching.jl:96,99
    integrate_vp!(particles, efield, lparams, IP, IC)  scripts/bump_on_tail_1_training.jl:97

and writes tests/golden/ref_julia_outputs.bin (raw little-endian Float64, layout below) plus
tests/golden/ref_julia_versions.txt.  `python -m pytest tests/test_reference_vectors.py` then compares the oracle
(every machine) and the GPU path (-m gpu) against these vectors and reports which convention of SURVEY.md App. A.4
(sign / gauge / cyclic tap shift of Phi, timing of A at a save) the real packages use.

Output layout (all Float64, column-major as Julia writes them):
    magic = 20261020.0, n, nh, nt, ns, nchi
    one update! at the initial positions with chi[2]:   rhs[nh] (NaN if the solver exposes no rhs field), phi[nh], a[n], W
    for each chi:   X[n*ns], V[n*ns], A[n*ns], Phi[nh*ns], W[ns], K[ns], M[ns]
=#
using ReducedBasisMethods
using ParticleMethods
using PoissonSolvers
using VlasovMethods
import Pkg

const ROOT = normpath(joinpath(@__DIR__, ".."))
const GOLD = joinpath(ROOT, "tests", "golden")

# ---- inputs
io = open(joinpath(GOLD, "ref_julia_inputs.bin"), "r")
n, nh, nt, ns, nchi = [read(io, Int64) for _ in 1:5]
L, dt, w0 = [read(io, Float64) for _ in 1:3]
chi = read!(io, Vector{Float64}(undef, nchi))
x0 = read!(io, Vector{Float64}(undef, n))
v0 = read!(io, Vector{Float64}(undef, n))
close(io)

p = 3                                                                  # scripts/bump_on_tail_1_training.jl:30
particles = ParticleList(reshape(copy(x0), 1, n), reshape(copy(v0), 1, n), fill(w0, 1, n))   # test/trainingset_tests.jl:17-21
poisson = PoissonSolverPBSplines(p, nh, L)
w = fill(w0, n)

out = open(joinpath(GOLD, "ref_julia_outputs.bin"), "w")
write(out, Float64[20261020.0, n, nh, nt, ns, nchi])

# ---- one update! / efield! at the initial positions (pins deposit, solve, gather, energy separately)
let efield = ScaledPoissonField(poisson, chi[2])
    x = copy(x0)
    E = zeros(n)
    VlasovMethods.update!(efield, x, w, 0.0)
    rhs = hasproperty(poisson, :rhs) ? Vector{Float64}(vec(getproperty(poisson, :rhs))) : fill(NaN, nh)
    phi = Vector{Float64}(vec(VlasovMethods.coefficients(efield)))
    VlasovMethods.efield!(efield, E, x)
    write(out, length(rhs) == nh ? rhs : fill(NaN, nh)); write(out, phi); write(out, E)
    write(out, Float64[VlasovMethods.energy(efield)])
end

# ---- the training loop of scripts/bump_on_tail_1_training.jl:87-109 for the three chi
IP = VPIntegratorParameters(dt, nt, ns, nh, n)                         # :67
IC = VPIntegratorCache(IP)                                             # :70
for c in chi
    efield = ScaledPoissonField(poisson, c)
    lparams = (χ = c,)
    integrate_vp!(particles, efield, lparams, IP, IC; save = true)
    for A in (IC.X, IC.V, IC.A, IC.Φ, IC.W, IC.K, IC.M)
        write(out, Vector{Float64}(vec(A)))
    end
end
close(out)

open(joinpath(GOLD, "ref_julia_versions.txt"), "w") do f
    println(f, "julia ", VERSION)
    for (_, pkg) in Pkg.dependencies()
        pkg.name in ("ReducedBasisMethods", "VlasovMethods", "PoissonSolvers", "ParticleMethods") &&
            println(f, pkg.name, " ", pkg.version, " ", something(pkg.git_revision, ""))
    end
end
println("wrote tests/golden/ref_julia_outputs.bin; now run: python -m pytest tests/test_reference_vectors.py -q")
