"""
    ReducedBasisMethodsB200

Julia host layer for librbm_b200.so (include/rbm_b200.h): keeps the API surface of
ReducedBasisMethods.jl for the offline hot path and forwards the compute through `ccall`.
Nothing here computes; every method is a declarative wrapper (no CUDA.jl, no CPU fallback).

NOTE: this file could not be executed where it was written (no Julia toolchain in the build
image); the same entry points are exercised through ctypes by tests/ (Python mirror `rbm_b200`).

Usage in the unmodified training script (scripts/bump_on_tail_1_training.jl):

    using ReducedBasisMethodsB200
    efield = B200PoissonField(poisson, lparams.χ)                       # instead of ScaledPoissonField (:94)
    integrate_vp_b200!(particles, efield, lparams, IP, IC; save=true)   # instead of integrate_vp!      (:97)

or, replacing the whole sample loop (:87-109) by one call that advances all samples in lock-step:

    integrate_vp_b200!(particles, poisson, pspace.samples.χ, TS.integrator, TS.snapshots)

and in scripts/bump_on_tail_2_projections.jl:26

    rb = ReducedBasis(CotangentLiftEVDB200(), ts; particle_tol, field_tol)
"""
module ReducedBasisMethodsB200

using Libdl
using ParticleMethods: ParticleList
using PoissonSolvers: PoissonSolverPBSplines
using ReducedBasisMethods
using ReducedBasisMethods: IntegratorParameters, Snapshots, TrainingSet, ReducedBasis, ReductionAlgorithm
import VlasovMethods
using VlasovMethods: ElectricField

export B200PoissonField, integrate_vp_b200!, reduced_integrate_vp_b200!, CotangentLiftEVDB200,
       get_PODBasis_cotangentLiftEVD_b200, get_PODBasis_EVD_b200, get_DEIM_interpolation_matrix_b200,
       rbm_context, rbm_comm_attach!, draw_accept_reject_b200, get_regression_αβ_b200, rbm_pin!, rbm_unpin!

const librbm = get(ENV, "RBM_B200_LIB", joinpath(@__DIR__, "..", "..", "..", "lib", "librbm_b200.so"))

# ----------------------------------------------------------------------------- context

mutable struct Context
    h::Ptr{Cvoid}
end

function check(ctx::Ptr{Cvoid}, rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:rbm_last_error, librbm), Cstring, (Ptr{Cvoid},), ctx))
    error("librbm_b200 error $rc: $msg")            # the reference signals errors with @assert / throw
end

function Context(device::Integer = parse(Int, get(ENV, "LOCAL_RANK", "0")))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:rbm_init, librbm), Cint, (Cint, Ref{Ptr{Cvoid}}), device, out)
    check(C_NULL, rc)
    ctx = Context(out[])
    finalizer(c -> ccall((:rbm_finalize, librbm), Cvoid, (Ptr{Cvoid},), c.h), ctx)
    ctx
end

const _ctx = Ref{Union{Nothing,Context}}(nothing)
rbm_context() = (_ctx[] === nothing && (_ctx[] = Context()); _ctx[]::Context)

"Attach an NCCL communicator: `id` is the 128-byte id created on rank 0 by `rbm_comm_unique_id` and broadcast by the host (MPI.jl, a file, ...)."
function rbm_comm_attach!(ctx::Context, nranks::Integer, rank::Integer, id::Vector{UInt8})
    @assert length(id) == 128
    check(ctx.h, ccall((:rbm_comm_init, librbm), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctx.h, nranks, rank, id))
end

function rbm_comm_unique_id()
    id = zeros(UInt8, 128)
    check(C_NULL, ccall((:rbm_comm_unique_id, librbm), Cint, (Ptr{UInt8},), id))
    id
end

"""
    rbm_pin!(A) / rbm_unpin!(A)

Page-lock a Julia `Array{Float64}` in place (rbm_host_register / rbm_host_unregister).  Julia arrays are pageable;
pinning `SS.X, SS.V, SS.A` once after `Snapshots(...)` lets `integrate_vp_b200!` stream every saved slot to the host at
pinned-memory speed, overlapped with the time loop.  Unpin before the array can be freed by the GC.
"""
function rbm_pin!(A::Array{Float64}; ctx::Context = rbm_context())
    check(ctx.h, ccall((:rbm_host_register, librbm), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64), ctx.h, A, sizeof(A)))
    A
end
function rbm_unpin!(A::Array{Float64}; ctx::Context = rbm_context())
    check(ctx.h, ccall((:rbm_host_unregister, librbm), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), ctx.h, A))
    A
end

# ----------------------------------------------------------------------------- PIC handle

mutable struct PIC
    h::Ptr{Cvoid}
    ctx::Context
    np::Int
    nh::Int
end

function PIC(np::Integer, poisson::PoissonSolverPBSplines; ctx::Context = rbm_context())
    out = Ref{Ptr{Cvoid}}(C_NULL)
    nh = length(poisson)                                                # src/trainingset.jl:21
    rc = ccall((:rbm_pic_create, librbm), Cint, (Ptr{Cvoid}, Int64, Cint, Cint, Cdouble, Ref{Ptr{Cvoid}}),
               ctx.h, np, nh, poisson.p, poisson.L, out)               # fields p, L: src/particles/poisson.jl:7-8
    check(ctx.h, rc)
    pic = PIC(out[], ctx, np, nh)
    finalizer(p -> ccall((:rbm_pic_destroy, librbm), Cvoid, (Ptr{Cvoid},), p.h), pic)
    pic
end

# Handles of the INTEGRATOR entry points are cached per (np, p, nh, L): they hold the particle state and the persistent
# workspaces of rbm_pic_run, and every call re-loads the particles.  A field object never uses this cache: each
# B200PoissonField owns its handle, because the handle holds that field's state (phi, chi, W) between update! and
# efield!/energy/coefficients -- two fields with different chi must not share it.
const _pics = Dict{Tuple{Int,Int,Int,Float64},PIC}()
pic_for(np, poisson) = get!(() -> PIC(np, poisson), _pics, (np, poisson.p, length(poisson), poisson.L))

# a contiguous Vector{Float64} is passed as it is; anything else (a row of ParticleList.list, a reshaped view) is copied once
_dense(a::Vector{Float64}) = a
_dense(a::AbstractArray) = (v = vec(a); v isa Vector{Float64} ? v : collect(Float64, v))

"""
    set_conventions!(pic; phi_sign = 1, tap_shift = 0, gauge = 0.0, a_at_save = :kick)

Output conventions that nothing in the reference checkout pins (sign / cyclic tap shift / gauge of the stored Φ, and whether
`IC.A` at a save is the kick's acceleration or the post-`update!` field).  `tools/reference_vectors.jl` +
`tests/test_reference_vectors.py` report the values of the real packages; set them here once.
"""
function set_conventions!(pic::PIC; phi_sign::Real = 1, tap_shift::Integer = 0, gauge::Real = 0.0, a_at_save::Symbol = :kick)
    a_at_save in (:kick, :post_update) || throw(ArgumentError("a_at_save must be :kick or :post_update"))
    check(pic.ctx.h, ccall((:rbm_pic_set_conventions, librbm), Cint, (Ptr{Cvoid}, Cdouble, Cint, Cdouble, Cint),
                           pic.h, Float64(phi_sign), Cint(tap_shift), Float64(gauge), Cint(a_at_save == :post_update)))
end

# ParticleList stores x, v, w as rows of one stacked array (array-of-structs in memory);
# the C ABI wants contiguous vectors, so copy once (np doubles each).
function set_particles!(pic::PIC, P::ParticleList)
    x = collect(vec(P.x)); v = collect(vec(P.v)); w = collect(vec(P.w))
    @assert length(x) == pic.np
    if all(==(w[1]), w)
        rc = ccall((:rbm_pic_set_particles, librbm), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble),
                   pic.h, x, v, C_NULL, w[1])
    else
        rc = ccall((:rbm_pic_set_particles, librbm), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble),
                   pic.h, x, v, w, 0.0)
    end
    check(pic.ctx.h, rc)
end

_ptr(a::Nothing) = Ptr{Cdouble}(C_NULL)
_ptr(a::AbstractArray{Float64}) = pointer(a)

function run!(pic::PIC, χ::Vector{Float64}, dt, nt, ns; X=nothing, V=nothing, A=nothing, Φ=nothing, W=nothing, K=nothing, M=nothing)
    GC.@preserve X V A Φ W K M begin
        rc = ccall((:rbm_pic_run, librbm), Cint,
                   (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cdouble, Cint, Cint,
                    Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                   pic.h, χ, length(χ), dt, nt, ns, _ptr(X), _ptr(V), _ptr(A), _ptr(Φ), _ptr(W), _ptr(K), _ptr(M))
    end
    check(pic.ctx.h, rc)
end

# ----------------------------------------------------------------------------- ElectricField protocol
# src/particles/electric_field.jl:22-35; used at src/particles/time_marching.jl:95-99,140

struct B200PoissonField{PT <: PoissonSolverPBSplines} <: ElectricField
    poisson::PT
    χ::Float64
    pic::PIC
end

# one handle per field object (np = 1: the protocol calls carry their own particle count)
B200PoissonField(poisson::PoissonSolverPBSplines, χ::Real; ctx::Context = rbm_context()) =
    B200PoissonField(poisson, Float64(χ), PIC(1, poisson; ctx = ctx))

function VlasovMethods.update!(f::B200PoissonField, x::AbstractArray, w::AbstractArray, t)
    xx = _dense(x); ww = _dense(w)
    length(xx) == length(ww) || throw(DimensionMismatch())
    GC.@preserve xx ww check(f.pic.ctx.h, ccall((:rbm_field_update, librbm), Cint,
                             (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Int64, Cdouble),
                             f.pic.h, xx, ww, 0.0, length(xx), f.χ))
end

function VlasovMethods.efield!(f::B200PoissonField, E::AbstractArray, x::AbstractArray)
    xx = _dense(x)
    length(E) == length(xx) || throw(DimensionMismatch())
    if E isa Array{Float64}            # written in place, no temporary
        GC.@preserve xx E check(f.pic.ctx.h, ccall((:rbm_field_eval, librbm), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Ptr{Cdouble}),
                                  f.pic.h, xx, length(xx), E))
    else
        ee = Vector{Float64}(undef, length(xx))
        check(f.pic.ctx.h, ccall((:rbm_field_eval, librbm), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Ptr{Cdouble}),
                                 f.pic.h, xx, length(xx), ee))
        E .= reshape(ee, size(E))
    end
    E
end

function VlasovMethods.energy(f::B200PoissonField)
    W = Ref{Cdouble}(0.0)
    check(f.pic.ctx.h, ccall((:rbm_field_energy, librbm), Cint, (Ptr{Cvoid}, Ref{Cdouble}), f.pic.h, W))
    W[]
end

function VlasovMethods.coefficients(f::B200PoissonField)
    ϕ = Vector{Float64}(undef, length(f.poisson))
    check(f.pic.ctx.h, ccall((:rbm_field_coefficients, librbm), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), f.pic.h, ϕ))
    ϕ
end

(f::B200PoissonField)(E, x, w, t) = (VlasovMethods.update!(f, x, w, t); VlasovMethods.efield!(f, E, x))

# ----------------------------------------------------------------------------- integrate_vp!

"Per-sample drop-in for `integrate_vp!(particles, efield, lparams, IP, IC; save)` (scripts/bump_on_tail_1_training.jl:97)."
function integrate_vp_b200!(P::ParticleList, efield::B200PoissonField, lparams::NamedTuple, IP, IC; save = true)
    pic = pic_for(length(P), efield.poisson)
    set_particles!(pic, P)
    if save
        run!(pic, [efield.χ], IP.dt, IP.nₜ, IP.nₛ; X = IC.X, V = IC.V, A = IC.A, Φ = IC.Φ, W = IC.W, K = IC.K, M = IC.M)
    else
        run!(pic, [efield.χ], IP.dt, IP.nₜ, 2)
    end
    IC
end

"All samples at once, straight into `Snapshots` (column-major (nd, np, ns, nparam) == the device layout)."
function integrate_vp_b200!(P::ParticleList, poisson::PoissonSolverPBSplines, χ::AbstractVector, IP::IntegratorParameters, SS::Snapshots;
                            pin::Bool = false)
    @assert IP.nparam == length(χ) && IP.nₚ == length(P) && IP.nₕ == length(poisson)
    # the library sees raw addresses: refuse containers that do not have exactly the expected extent
    for A in (SS.X, SS.V, SS.A)
        length(A) == IP.nₚ * IP.nₛ * IP.nparam || throw(DimensionMismatch())
    end
    length(SS.Φ) == IP.nₕ * IP.nₛ * IP.nparam || throw(DimensionMismatch())
    for A in (SS.W, SS.K, SS.M)
        length(A) == IP.nₛ * IP.nparam || throw(DimensionMismatch())
    end
    pic = pic_for(length(P), poisson)
    set_particles!(pic, P)
    pin && foreach(A -> rbm_pin!(A; ctx = pic.ctx), (SS.X, SS.V, SS.A))   # pageable Julia arrays -> pinned-speed slot copies
    try
        run!(pic, collect(Float64, χ), IP.dt, IP.nₜ, IP.nₛ; X = SS.X, V = SS.V, A = SS.A, Φ = SS.Φ, W = SS.W, K = SS.K, M = SS.M)
    finally
        pin && foreach(A -> rbm_unpin!(A; ctx = pic.ctx), (SS.X, SS.V, SS.A))
    end
    SS
end

# ----------------------------------------------------------------------------- POD / DEIM

function _pod(blocks::Vector{<:AbstractMatrix{Float64}}, k::Integer, tolerance::Real; ctx = rbm_context())
    nrows = size(blocks[1], 1)
    all(size(b, 1) == nrows for b in blocks) || throw(DimensionMismatch())
    mats = [Matrix{Float64}(b) for b in blocks]                      # contiguous column-major copies of views
    ptrs = [pointer(m) for m in mats]; ncols = Int64[size(m, 2) for m in mats]; lds = Int64[nrows for _ in mats]
    C = sum(ncols)
    Λ = Vector{Float64}(undef, C); kout = Ref{Cint}(0)
    local Ψ
    GC.@preserve mats begin
        # factor (Gram + eigen + sorteigen + rank, src/algorithms/evd.jl:32-46), allocate Ψ with exactly k columns, then
        # form the basis (:49) from the explicit decomposition handle; nothing is cached inside the library
        f = Ref{Ptr{Cvoid}}(C_NULL)
        check(ctx.h, ccall((:rbm_pod_factor, librbm), Cint,
              (Ptr{Cvoid}, Ptr{Ptr{Cdouble}}, Ptr{Int64}, Ptr{Int64}, Cint, Int64, Cint, Cdouble, Ptr{Cdouble}, Ref{Cint}, Ref{Ptr{Cvoid}}),
              ctx.h, ptrs, ncols, lds, length(mats), nrows, k, tolerance, Λ, kout, f))
        try
            k = Int(kout[])
            Ψ = Matrix{Float64}(undef, nrows, k)
            check(ctx.h, ccall((:rbm_pod_basis, librbm), Cint,
                  (Ptr{Cvoid}, Ptr{Ptr{Cdouble}}, Ptr{Int64}, Ptr{Int64}, Cint, Int64, Cint, Ptr{Cdouble}, Int64),
                  f[], ptrs, ncols, lds, length(mats), nrows, k, Ψ, nrows))
        finally
            ccall((:rbm_evd_destroy, librbm), Cvoid, (Ptr{Cvoid},), f[])
        end
    end
    Ψ, Λ
end

"`get_PODBasis_cotangentLiftEVD(X, V; k, tolerance)` (src/algorithms/evd.jl:26-55) on the GPU; [X V] is never concatenated."
function get_PODBasis_cotangentLiftEVD_b200(X, V; k = 0, tolerance = 1e-8)
    @assert size(X) == size(V)
    _pod([X, V], k, tolerance)
end

"`get_PODBasis_EVD(S; k, tolerance)` (src/algorithms/evd.jl:63-87) on the GPU."
get_PODBasis_EVD_b200(S; k = 0, tolerance = 1e-4) = _pod([S], k, tolerance)

"`get_DEIM_interpolation_matrix(Ψ)` (src/algorithms/deim.jl:6-21): dense N×k 0/1 matrix, built from the selected rows."
function get_DEIM_interpolation_matrix_b200(Ψ::AbstractMatrix{Float64}; ctx = rbm_context())
    N, k = size(Ψ); P = Matrix{Float64}(Ψ); idx = Vector{Int64}(undef, k)
    check(ctx.h, ccall((:rbm_deim, librbm), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Int64, Cint, Ptr{Int64}), ctx.h, P, N, N, k, idx))
    Π = zeros(N, k)
    for i in 1:k
        Π[idx[i] + 1, i] = 1.0                                       # 0-based rows cross the boundary
    end
    Π
end

"Dispatch tag next to `CotangentLiftEVD` (src/reducedbasis.jl:2-7)."
struct CotangentLiftEVDB200 <: ReductionAlgorithm end

function ReducedBasisMethods.ReducedBasis(alg::CotangentLiftEVDB200, ts::TrainingSet; particle_tol = 1e-8, field_tol = 1e-4)
    IP = ts.integrator                                               # src/algorithms/evd.jl:119-137
    X = reshape(ts.snapshots.X, (IP.nₚ, IP.nₛ * IP.nparam))
    V = reshape(ts.snapshots.V, (IP.nₚ, IP.nₛ * IP.nparam))
    E = reshape(ts.snapshots.A, (IP.nₚ, IP.nₛ * IP.nparam))
    Ψₚ, Λₚ = get_PODBasis_cotangentLiftEVD_b200(X, V; tolerance = particle_tol)
    Ψₑ, Λₑ = get_PODBasis_EVD_b200(E; tolerance = field_tol)
    Πₑ = get_DEIM_interpolation_matrix_b200(Ψₑ)
    ReducedBasis(alg, ts.parameters, ts.paramspace, ts.initconds, ts.integrator, ts.poisson, Λₚ, Ψₚ, Λₑ, Ψₑ, Πₑ)
end

# ----------------------------------------------------------------------------- reduced online model with DEIM

"""
    reduced_integrate_vp_b200!(P₀, poisson, Ψₚ, Ψₑ, Πₑ, χ, IP, RSS)

`DEIMElectricField` + `reduced_integrate_vp` (src/particles/electric_field.jl:39-43, src/particles/time_marching.jl:108-150)
for ALL test samples `χ` in one call; fills `RSS.X, V, Φ, W, K, M` (like the reference's `save_solution`, `RSS.A` is not written).
"""
function reduced_integrate_vp_b200!(P::ParticleList, poisson::PoissonSolverPBSplines, Ψₚ::Matrix{Float64}, Ψₑ::Matrix{Float64},
                                    Πₑ::AbstractMatrix, χ::AbstractVector, IP::IntegratorParameters, RSS::Snapshots)
    pic = pic_for(length(P), poisson)
    set_particles!(pic, P)
    N, kp = size(Ψₚ); ke = size(Ψₑ, 2)
    idx = Int64[findfirst(!iszero, view(Πₑ, :, j)) - 1 for j in 1:ke]          # 0-based DEIM rows
    rm = Ref{Ptr{Cvoid}}(C_NULL)
    check(pic.ctx.h, ccall((:rbm_reduced_create, librbm), Cint,
          (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Cint, Ptr{Cdouble}, Int64, Cint, Ptr{Int64}, Ref{Ptr{Cvoid}}),
          pic.h, Ψₚ, N, kp, Ψₑ, N, ke, idx, rm))
    χv = collect(Float64, χ)
    rc = ccall((:rbm_reduced_run, librbm), Cint,
               (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cdouble, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
                Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
               rm[], χv, length(χv), IP.dt, IP.nₜ, IP.nₛ, C_NULL, C_NULL, RSS.X, RSS.V, RSS.Φ, RSS.W, RSS.K, RSS.M)
    ccall((:rbm_reduced_destroy, librbm), Cvoid, (Ptr{Cvoid},), rm[])
    check(pic.ctx.h, rc)
    RSS
end

# ----------------------------------------------------------------------------- sampler and regression on the device

"""
    draw_accept_reject_b200(np, params; seed = 1234, first = 0, total = np) -> ParticleList

Device draw of the bump-on-tail initial condition (replaces `BumpOnTail.draw_accept_reject(np, params)` +
`Random.seed!(1234)`, scripts/bump_on_tail_1_training.jl:76-79).  Counter-based (Philox4x32-10): NOT the
MersenneTwister stream of the reference; particle `first + i` depends only on `(seed, first + i)`.
"""
function draw_accept_reject_b200(np::Integer, params::NamedTuple; seed::Integer = 1234, first::Integer = 0,
                                 total::Integer = np, ctx::Context = rbm_context())
    x = Vector{Float64}(undef, np); v = similar(x); w = similar(x)
    L = 2π / params.κ
    rej = Ref{Int64}(0)
    check(ctx.h, ccall((:rbm_sample_bump_on_tail, librbm), Cint,
          (Ptr{Cvoid}, Int64, Int64, UInt64, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble,
           Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Int64}),
          ctx.h, np, first, UInt64(seed), params.κ, params.ε, params.a, params.v₀, params.σ, L / total, x, v, w, rej))
    ParticleList(reshape(x, 1, :), reshape(v, 1, :), reshape(w, 1, :))    # (nd x np, nd x np, 1 x np) as in test/trainingset_tests.jl:17-21
end

"""
    get_regression_αβ_b200(t, W, n, inds)

`get_regression_αβ(t, W, n, inds)` of src/regression.jl:14-24 (W is nₛ × nparam, `inds` 1-based positions in the
list of local maxima, as in the reference).
"""
function get_regression_αβ_b200(t::AbstractVector, W::Matrix{Float64}, n::Integer, inds; ctx::Context = rbm_context())
    ns, nparam = size(W)
    α = zeros(nparam); β = zeros(nparam)
    i0 = Cint[i - 1 for i in inds]
    rc = ccall((:rbm_growth_rate, librbm), Cint,
               (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Int64, Int64, Cint, Ptr{Cint}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}),
               ctx.h, collect(Float64, t), W, ns, nparam, ns, n, i0, length(i0), α, β, C_NULL)
    rc == -1 && occursin("BoundsError", unsafe_string(ccall((:rbm_last_error, librbm), Cstring, (Ptr{Cvoid},), ctx.h))) &&
        throw(BoundsError(W, inds))
    check(ctx.h, rc)
    α, β
end

end # module
