# VPMB200.jl — Julia host shim over libvpm_b200.so (include/vpm_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia toolchain (SURVEY F3).  The same C ABI
# is exercised end to end from Python (vlasovparticlemethods.jl_b200/api.py); this file is the binding a
# VlasovMethods.jl maintainer would add.  It defines device-backed subtypes of the package's own abstract
# types and adds methods to the package's generic functions, so scripts/*.jl change only in their
# constructor lines (see INTEGRATION.md).
module VPMB200

using VlasovMethods
import VlasovMethods: projection!, projection, run!, initialize!, DistributionFunction,
                      SplittingMethod, GeometricIntegrator, VlasovPoisson,
                      LenardBernstein, ConservativeLenardBernstein, BumpOnTail, DoubleMaxwellian,
                      UniformDistribution, ShiftedUniformDistribution, ShiftedNormalV,
                      LB_rhs!, CLB_rhs!, LB_rhs_GI!, CLB_rhs_GI!, s_advection!, s_acceleration!,
                      compute_f_densities, compute_df_densities, projection_density, projection_momentum, projection_energy
using BSplineKit: Derivative
import PoissonSolvers   # a direct dependency of VlasovMethods (Project.toml:42); the model calls PoissonSolvers.update!

const libvpm = get(ENV, "LIBVPM_B200", "libvpm_b200.so")

struct VPMError <: Exception
    code::Cint
    msg::String
end

@inline function check(rc::Cint)
    rc == 0 && return nothing
    throw(VPMError(rc, unsafe_string(ccall((:vpm_last_error, libvpm), Cstring, ()))))
end

# ---------------------------------------------------------------------------------- context
mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = 0; stream::Ptr{Cvoid} = C_NULL)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:vpm_ctx_create, libvpm), Cint, (Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), device, stream, r))
        finalizer(c -> ccall((:vpm_ctx_destroy, libvpm), Cint, (Ptr{Cvoid},), c.h), new(r[]))
    end
end
const DEFAULT = Ref{Union{Nothing,Context}}(nothing)
context() = something(DEFAULT[], (DEFAULT[] = Context(0)))

# ---------------------------------------------------------------------------------- distributions
# replaces ParticleDistribution (src/distributions/particle_distribution.jl:2-20): SoA on the device
mutable struct DeviceParticleDistribution <: DistributionFunction{1,1}
    h::Ptr{Cvoid}
    n::Int
    ctx::Context
    function DeviceParticleDistribution(npart::Integer; ctx = context())
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:vpm_particles_create, libvpm), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{Cvoid}}), ctx.h, npart, r))
        finalizer(p -> ccall((:vpm_particles_destroy, libvpm), Cint, (Ptr{Cvoid},), p.h), new(r[], npart, ctx))
    end
end
Base.size(d::DeviceParticleDistribution) = (d.n,)
Base.length(d::DeviceParticleDistribution) = d.n

# host <-> device with the reference's own (x;v;w) x N matrix (ld = 3) or the integrator state z (ld = 2)
function upload!(d::DeviceParticleDistribution, z::Matrix{Float64})
    check(ccall((:vpm_particles_upload_aos, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint), d.h, z, size(z, 1)))
    d
end
function download!(z::Matrix{Float64}, d::DeviceParticleDistribution)
    check(ccall((:vpm_particles_download_aos, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint), d.h, z, size(z, 1)))
    z
end
DeviceParticleDistribution(dist::VlasovMethods.ParticleDistribution; kw...) =
    upload!(DeviceParticleDistribution(length(dist.particles); kw...), Matrix(dist.particles.list))

# replaces SplineDistribution (src/distributions/spline_distribution.jl:23-36)
mutable struct DeviceSplineDistribution <: DistributionFunction{1,1}
    h::Ptr{Cvoid}
    n::Int
    ctx::Context
    function DeviceSplineDistribution(nknots::Integer, order::Integer, domain::Tuple, bc::Symbol = :Dirichlet; ctx = context())
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:vpm_vspace_create, libvpm), Cint, (Ptr{Cvoid}, Float64, Float64, Cint, Cint, Cint, Ref{Ptr{Cvoid}}),
                    ctx.h, domain[1], domain[2], nknots, order, bc == :Dirichlet, r))
        n = ccall((:vpm_vspace_size, libvpm), Cint, (Ptr{Cvoid},), r[])
        finalizer(s -> ccall((:vpm_vspace_destroy, libvpm), Cint, (Ptr{Cvoid},), s.h), new(r[], n, ctx))
    end
end
Base.size(s::DeviceSplineDistribution) = (s.n,)
function coefficients(s::DeviceSplineDistribution)
    c = Vector{Float64}(undef, s.n)
    check(ccall((:vpm_vspace_get, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), s.h, C_NULL, c))
    c
end

# replaces Potential(PeriodicBasisBSplineKit(domain, order, n)) (scripts/vlasov_poisson.jl:21)
mutable struct DevicePotential
    h::Ptr{Cvoid}
    n::Int
    ctx::Context
    function DevicePotential(domain::Tuple, order::Integer, n::Integer; ctx = context())
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:vpm_xspace_create, libvpm), Cint, (Ptr{Cvoid}, Float64, Float64, Cint, Cint, Ref{Ptr{Cvoid}}),
                    ctx.h, domain[1], domain[2], order, n, r))
        finalizer(p -> ccall((:vpm_xspace_destroy, libvpm), Cint, (Ptr{Cvoid},), p.h), new(r[], n, ctx))
    end
end
function Base.getproperty(p::DevicePotential, s::Symbol)
    if s === :rhs || s === :coefficients
        out = Vector{Float64}(undef, getfield(p, :n))
        a, b = s === :rhs ? (out, C_NULL) : (C_NULL, out)
        check(ccall((:vpm_xspace_get, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), getfield(p, :h), a, b))
        return out
    end
    getfield(p, s)
end

# Models.  VlasovPoisson{XD,VD,DT,PT<:Potential} (src/models/vlasov_poisson.jl:1-8) constrains its potential to
# PoissonSolvers' own Potential type and CollisionEntropy builds a SplineDistributionCache that only accepts the host
# SplineDistribution (src/entropies/collision_entropy.jl:1-10, src/distributions/spline_distribution.jl:41-51), so the
# device types get their own small holders, reached through more specific methods of the SAME constructor names: the
# scripts keep writing VlasovPoisson(dist, potential) and CollisionEntropy(sdist).
struct DeviceVlasovPoisson <: VlasovMethods.VlasovModel
    distribution::DeviceParticleDistribution
    potential::DevicePotential
end
VlasovMethods.VlasovPoisson(dist::DeviceParticleDistribution, potential::DevicePotential) = DeviceVlasovPoisson(dist, potential)

struct DeviceCollisionEntropy <: VlasovMethods.Entropy
    dist::DeviceSplineDistribution
end
VlasovMethods.CollisionEntropy(dist::DeviceSplineDistribution) = DeviceCollisionEntropy(dist)
# LenardBernstein(dist, ent; ν) / ConservativeLenardBernstein(dist, ent; ν) take any DistributionFunction{1,1} and any
# Entropy (src/models/lenard_bernstein.jl:1-9), so they work with the device types as they are.

# compute_entropy!(entropy, dist): a TODO upstream (src/entropies/collision_entropy.jl:12-15) -- NON-REFERENCE diagnostic:
# S = -sum_p w_p ln max(f_s(v_p), f_floor) with f_s the spline projection of the particles (vpm_entropy_v)
const ENTROPY_FLOOR = 1e-14
function compute_entropy!(entropy::DeviceCollisionEntropy, dist::DeviceParticleDistribution; f_floor::Float64 = ENTROPY_FLOOR)
    _, v, w = ptrs(dist)
    check(ccall((:vpm_project_v, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}),
                entropy.dist.h, v, w, dist.n, C_NULL))
    S = Ref{Float64}(); nf = Ref{Float64}()
    check(ccall((:vpm_entropy_v, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Float64, Ref{Float64}, Ref{Float64}),
                entropy.dist.h, C_NULL, v, w, dist.n, f_floor, S, nf))
    S[]
end

# every reference sampler gives equal weights: declaring it lets the steppers skip the w[] stream
set_uniform_weight!(d::DeviceParticleDistribution, w::Real) =
    (check(ccall((:vpm_particles_set_uniform_weight, libvpm), Cint, (Ptr{Cvoid}, Float64), d.h, w)); d)

# ---------------------------------------------------------------------------------- multi-GPU (one process per GPU)
# Fused peer-memory all-reduce of the coefficient vector inside the field kernels (NVLink).  `allgather` is any
# host-side all-gather of 64-byte blobs in rank order (e.g. MPI.Allgather); NCCL (`vpm_comm_init`) is the fallback.
function attach_peers!(ctx::Context, nranks::Integer, rank::Integer, allgather::Function)
    h = Vector{UInt8}(undef, 64)
    check(ccall((:vpm_p2p_prepare, libvpm), Cint, (Ptr{Cvoid}, Ptr{UInt8}), ctx.h, h))
    all = allgather(h)::Vector{UInt8}                       # 64 * nranks bytes
    check(ccall((:vpm_p2p_attach, libvpm), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctx.h, nranks, rank, all))
    ctx
end

# ---------------------------------------------------------------------------------- seam functions
# read-only device pointers (the writable form, vpm_particles_ptrs, ends a uniform-weight declaration)
function ptrs(d::DeviceParticleDistribution)
    x = Ref{Ptr{Float64}}(); v = Ref{Ptr{Float64}}(); w = Ref{Ptr{Float64}}()
    check(ccall((:vpm_particles_ptrs_const, libvpm), Cint, (Ptr{Cvoid}, Ref{Ptr{Float64}}, Ref{Ptr{Float64}}, Ref{Ptr{Float64}}), d.h, x, v, w))
    x[], v[], w[]
end

# projection!(potential, distribution): src/projections/potential.jl:2-22
function projection!(potential::DevicePotential, distribution::DeviceParticleDistribution)
    x, _, w = ptrs(distribution)
    check(ccall((:vpm_deposit_x, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}),
                potential.h, x, w, distribution.n, C_NULL))
    potential
end

# PoissonSolvers.update!(potential): call site src/models/vlasov_poisson.jl:14.  A METHOD OF PoissonSolvers' OWN GENERIC
# FUNCTION, so that update_potential!(model) -- which calls PoissonSolvers.update!(model.potential) -- dispatches here
# for a DevicePotential instead of raising a MethodError.
PoissonSolvers.update!(potential::DevicePotential) =
    (check(ccall((:vpm_poisson_solve, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), potential.h, C_NULL, C_NULL)); potential)
const update! = PoissonSolvers.update!

# update_potential!(model): src/models/vlasov_poisson.jl:12-15
VlasovMethods.update_potential!(model::DeviceVlasovPoisson) =
    check(ccall((:vpm_update_potential, libvpm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
                model.potential.h, model.distribution.h, C_NULL, C_NULL))

# projection(velocities, dist, final_dist): src/projections/distribution.jl:35-55 (velocities on the host)
function projection(velocities::AbstractVector{Float64}, dist::DeviceParticleDistribution, final_dist::DeviceSplineDistribution)
    ctx = dist.ctx
    dv = Ref{Ptr{Float64}}()
    check(ccall((:vpm_dev_alloc, libvpm), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{Float64}}), ctx.h, length(velocities), dv))
    try
        check(ccall((:vpm_memcpy_h2d, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), ctx.h, dv[], velocities, length(velocities)))
        _, _, w = ptrs(dist)
        check(ccall((:vpm_project_v, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}),
                    final_dist.h, dv[], w, dist.n, C_NULL))
    finally
        ccall((:vpm_dev_free, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx.h, dv[])
    end
    DeviceSpline(final_dist, false)     # aliases final_dist's coefficients, as Spline(basis, coefficients) does upstream
end

# with_device_vector(f, ctx, host): f(device pointer) on a temporary device copy of a host vector
function with_device_vector(f::Function, ctx::Context, host::AbstractVector{Float64})
    d = Ref{Ptr{Float64}}()
    check(ccall((:vpm_dev_alloc, libvpm), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{Float64}}), ctx.h, length(host), d))
    try
        check(ccall((:vpm_memcpy_h2d, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), ctx.h, d[], host, length(host)))
        return f(d[])
    finally
        ccall((:vpm_dev_free, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx.h, d[])
    end
end

# fs.(v) and (Derivative(1) * fs).(v): src/models/lenard_bernstein.jl:26-28, src/projections/density.jl:45-47
struct DeviceSpline
    sdist::DeviceSplineDistribution
    derivative::Bool
end
Base.:*(::Derivative{1}, s::DeviceSpline) = DeviceSpline(s.sdist, true)
function (s::DeviceSpline)(v::AbstractVector{Float64})
    n, ctx = length(v), s.sdist.ctx
    out = Vector{Float64}(undef, n)
    with_device_vector(ctx, collect(v)) do dv
        with_device_vector(ctx, out) do dout
            f, df = s.derivative ? (Ptr{Float64}(C_NULL), dout) : (dout, Ptr{Float64}(C_NULL))
            check(ccall((:vpm_gather_v, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}),
                        s.sdist.h, C_NULL, dv, n, f, df))
            check(ccall((:vpm_memcpy_d2h, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), ctx.h, out, dout, n))
        end
    end
    out
end
(s::DeviceSpline)(v::Real) = s([Float64(v)])[1]

# compute_f_densities / compute_df_densities / projection_*: src/projections/density.jl:6-52 (unweighted sums, one pass)
function moments(sdist::DeviceSplineDistribution, vp::AbstractVector{Float64})
    out5 = Vector{Float64}(undef, 5)
    with_device_vector(sdist.ctx, collect(vp)) do dv
        check(ccall((:vpm_moments, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}),
                    sdist.h, C_NULL, dv, length(vp), out5))
    end
    out5      # Σf, Σvf, Σv²f, Σf', Σvf'
end
compute_f_densities(sdist::DeviceSplineDistribution, vp) = (m = moments(sdist, vp); (m[1], m[2], m[3]))
compute_df_densities(sdist::DeviceSplineDistribution, vp) = (m = moments(sdist, vp); (m[4], m[5]))
projection_density(sdist::DeviceSplineDistribution, vp; isDerivative = false) = moments(sdist, vp)[isDerivative ? 4 : 1]
projection_momentum(sdist::DeviceSplineDistribution, vp; isDerivative = false) = moments(sdist, vp)[isDerivative ? 5 : 2]
projection_energy(sdist::DeviceSplineDistribution, vp) = moments(sdist, vp)[3]

# ϕ(x, Derivative(1)) for a host vector of positions: call sites src/models/vlasov_poisson.jl:27,48,65
function (ϕ::DevicePotential)(x::AbstractVector{Float64}, ::Derivative{1})
    n, ctx = length(x), getfield(ϕ, :ctx)
    coef = ϕ.coefficients
    out = Vector{Float64}(undef, n)
    with_device_vector(ctx, collect(x)) do dx
        with_device_vector(ctx, out) do dout
            check(ccall((:vpm_gather_x, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Cint, Ptr{Float64}),
                        getfield(ϕ, :h), coef, dx, n, 1, dout))
            check(ccall((:vpm_memcpy_d2h, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), ctx.h, out, dout, n))
        end
    end
    out
end
(ϕ::DevicePotential)(x::Real, d::Derivative{1}) = ϕ([Float64(x)], d)[1]

# The splitting flows with the reference's callback signature (z, t, z̄, t̄, params) on host matrices
# (src/models/vlasov_poisson.jl:53-67), for callers that keep GeometricIntegrators in charge of the time loop.
# Every call moves the state over PCIe: for parity, not for speed (use run! for that).
const DeviceVPParams = NamedTuple{(:ϕ, :model),<:Tuple{DevicePotential,Any}}
function s_advection!(z, t, z̄, t̄, params::DeviceVPParams)
    z[1, :] .= z̄[1, :] .+ (t - t̄) .* z̄[2, :]
    z[2, :] .= z̄[2, :]
    z
end
function s_acceleration!(z, t, z̄, t̄, params::DeviceVPParams)
    VlasovMethods.update_potential!(params.model)            # deposits from model.distribution (SURVEY F4)
    z[1, :] .= z̄[1, :]
    z[2, :] .= z̄[2, :] .- (t - t̄) .* params.ϕ(z̄[1, :], Derivative(1))
    z
end
# one whole Strang step of the host state through the device (the `e2e` path of bench.py)
function strang_step!(z::Matrix{Float64}, z̄::Matrix{Float64}, m; selfconsistent::Bool = false)
    check(ccall((:vpm_vp_strang_step_host, libvpm), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Cint),
                m.model.potential.h, m.model.distribution.h, z̄, z, m.tstep, m.χ, selfconsistent ? 0 : 1))
    z
end

# LB_rhs! / CLB_rhs!: src/models/lenard_bernstein.jl:20-30, lenard_bernstein_conservative.jl:24-36
function lb_rhs!(v̇::Vector{Float64}, v::Vector{Float64}, params, conservative::Bool)
    dist, sdist, ctx = params.idist, params.model.ent.dist, params.idist.ctx
    n = length(v)
    dv = Ref{Ptr{Float64}}(); dout = Ref{Ptr{Float64}}()
    check(ccall((:vpm_dev_alloc, libvpm), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{Float64}}), ctx.h, n, dv))
    check(ccall((:vpm_dev_alloc, libvpm), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{Float64}}), ctx.h, n, dout))
    try
        check(ccall((:vpm_memcpy_h2d, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), ctx.h, dv[], v, n))
        _, _, w = ptrs(dist)
        check(ccall((:vpm_lb_rhs, libvpm), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Float64, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                    sdist.h, dv[], w, n, params.ν, conservative, dout[], C_NULL, C_NULL))
        check(ccall((:vpm_memcpy_d2h, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), ctx.h, v̇, dout[], n))
    finally
        ccall((:vpm_dev_free, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx.h, dv[])
        ccall((:vpm_dev_free, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx.h, dout[])
    end
    v̇
end
# the reference's own signatures: DiffEq form (v̇, v, params, t) and GeometricIntegrators form (v, t, q, params)
const DeviceLBParams = NamedTuple{(:ν, :idist, :fdist, :model),<:Tuple{Any,DeviceParticleDistribution,Any,Any}}
# (AbstractVector{Float64} + the params type make these strictly more specific than the package's own methods,
#  LB_rhs!(v̇, v::AbstractArray{ST}, params, t) and CLB_rhs!(v̇, v::AbstractVector{ST}, params, t): no ambiguity)
LB_rhs!(v̇, v::AbstractVector{Float64}, params::DeviceLBParams, t) =
    (v̇ .= lb_rhs!(Vector{Float64}(undef, length(v)), collect(v), params, false))
CLB_rhs!(v̇, v::AbstractVector{Float64}, params::DeviceLBParams, t) =
    (v̇ .= lb_rhs!(Vector{Float64}(undef, length(v)), collect(v), params, true))
LB_rhs_GI!(v, t, q::AbstractVector{Float64}, params::DeviceLBParams) = LB_rhs!(v, q, params, t)
CLB_rhs_GI!(v, t, q::AbstractVector{Float64}, params::DeviceLBParams) = CLB_rhs!(v, q, params, t)

# ---------------------------------------------------------------------------------- whole-run drivers
# SplittingMethod(model, tspan, tstep) + run!: src/models/vlasov_poisson.jl:73-89, src/methods/splitting.jl:23-52
struct DeviceSplittingMethod{MT}
    model::MT
    tspan::Tuple{Float64,Float64}
    tstep::Float64
    field::Symbol          # :frozen == as shipped (SURVEY F4), :selfconsistent == legacy integrate_vp!
    χ::Float64
end
SplittingMethod(model::DeviceVlasovPoisson, tspan::Tuple, tstep::Real;
                field::Symbol = :frozen, χ::Real = 1.0) = DeviceSplittingMethod(model, Float64.(tspan), Float64(tstep), field, Float64(χ))

# run!(method, h5file) as upstream (src/methods/splitting.jl:23-52): the trajectory goes to dataset "z" of h5file with the
# reference's layout (nd, np, nt+1), chunk (nd, np, 1), written by the library itself while the next steps compute, so
# `z = h5read(h5file, "z")` in scripts/vlasov_poisson.jl:38 keeps working.  save_stride = k keeps every k-th step (plus the
# last); the times of the saved frames are in dataset "t".  run!(method) without a file keeps everything on the device.
function run!(m::DeviceSplittingMethod, h5file::Union{AbstractString,Nothing} = nothing; save_stride::Integer = 1, diag_mode::Integer = 1)
    nt = Int(abs(div(m.tspan[2] - m.tspan[1], m.tstep, RoundUp)))   # GeometricEquations' ntime
    diag = zeros(3, nt + 1)          # rows W, K, M (src/vlasov_poisson.jl:58-67)
    frames = Ref{Cint}(0)
    check(ccall((:vpm_vp_run, libvpm), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64, Cint, Cint, Cint, Cint, Cstring, Ptr{Float64}, Ref{Cint}),
                m.model.potential.h, m.model.distribution.h, m.tstep, m.χ, nt, m.field === :frozen ? 1 : 0, diag_mode,
                h5file === nothing ? 0 : save_stride, h5file === nothing ? C_NULL : h5file, diag, frames))
    m.model.distribution, diag
end

# GeometricIntegrator(model, tspan, tstep) + run! with RK438: src/models/lenard_bernstein.jl:68-84,
# lenard_bernstein_conservative.jl:88-104, src/methods/geometric_integrator.jl:12-44
struct DeviceRK438{MT}
    model::MT
    tspan::Tuple{Float64,Float64}
    tstep::Float64
end
GeometricIntegrator(model::Union{LenardBernstein{1,1,DeviceParticleDistribution},ConservativeLenardBernstein{1,1,DeviceParticleDistribution}},
                    tspan::Tuple, tstep::Real) = DeviceRK438(model, Float64.(tspan), Float64(tstep))

# run!(method, h5file): datasets "z" (np, nt+1) chunk (np, 1) and "t" (nt+1) as src/methods/geometric_integrator.jl:21-35
# entropy = true also returns the history of the (non-reference) collision entropy S(t_n), see compute_entropy!
function run!(m::DeviceRK438, h5file::Union{AbstractString,Nothing} = nothing; save_stride::Integer = 1, entropy::Bool = false,
              f_floor::Float64 = ENTROPY_FLOOR)
    nt = Int(abs(div(m.tspan[2] - m.tspan[1], m.tstep, RoundUp)))   # GeometricEquations' ntime
    diag = zeros(2, nt + 1)          # rows Σv, Σv² (scripts/lenard_bernstein_conservative.jl:49-50)
    frames = Ref{Cint}(0)
    sh = m.model.ent.dist.h
    check(ccall((:vpm_vspace_entropy_history, libvpm), Cint, (Ptr{Cvoid}, Cint, Float64), sh, entropy, f_floor))
    S = entropy ? zeros(nt + 1) : nothing
    try
        check(ccall((:vpm_lb_run, libvpm), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64, Float64, Cint, Cint, Cint, Cstring, Ptr{Float64}, Ref{Cint}),
                    sh, m.model.dist.h, m.model.ν, m.tstep, m.tspan[1], nt, m.model isa ConservativeLenardBernstein,
                    h5file === nothing ? 0 : save_stride, h5file === nothing ? C_NULL : h5file, diag, frames))
        entropy && check(ccall((:vpm_vspace_entropy_get, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint), sh, S, C_NULL, nt + 1))
    finally
        ccall((:vpm_vspace_entropy_history, libvpm), Cint, (Ptr{Cvoid}, Cint, Float64), sh, 0, f_floor)
    end
    entropy ? (m.model.dist, diag, S) : (m.model.dist, diag)
end

# ---------------------------------------------------------------------------------- initial conditions
# initialize!(dist, BumpOnTail()): src/examples/bumpontail.jl:43-75 (device-side, counter-based)
function initialize!(d::DeviceParticleDistribution, p::BumpOnTail; seed::UInt64 = 0x000000005EED0001, offset::Integer = 0, ntotal::Integer = d.n)
    check(ccall((:vpm_sample_bump_on_tail, libvpm), Cint, (Ptr{Cvoid}, Int64, Int64, UInt64, Float64, Float64, Float64, Float64, Float64),
                d.h, offset, ntotal, seed, p.ε, p.κ, p.α, p.σ, p.v₀))
    d
end
function initialize!(d::DeviceParticleDistribution, p::DoubleMaxwellian; seed::UInt64 = 0x000000005EED0001, offset::Integer = 0, ntotal::Integer = d.n)
    check(ccall((:vpm_sample_maxwellian, libvpm), Cint, (Ptr{Cvoid}, Int64, Int64, UInt64, Float64, Float64, Float64, Cint, Float64),
                d.h, offset, ntotal, seed, p.domain[1], p.domain[2], p.shift, 1, 1.0))
    d
end
# UniformDistribution / ShiftedUniformDistribution / ShiftedNormalV: src/examples/{uniform,shifteduniform,shiftednormalv}.jl
function initialize!(d::DeviceParticleDistribution, p::Union{UniformDistribution,ShiftedUniformDistribution}; seed::UInt64 = 0x000000005EED0001, offset::Integer = 0, ntotal::Integer = d.n)
    shift = p isa ShiftedUniformDistribution ? Float64(p.shift) : 0.0
    check(ccall((:vpm_sample_uniform, libvpm), Cint, (Ptr{Cvoid}, Int64, Int64, UInt64, Float64, Float64, Float64, Float64, Float64, Float64),
                d.h, offset, ntotal, seed, p.xdomain[1], p.xdomain[2], p.vdomain[1], p.vdomain[2], shift, 1.0))
    d
end
function initialize!(d::DeviceParticleDistribution, p::ShiftedNormalV; seed::UInt64 = 0x000000005EED0001, offset::Integer = 0, ntotal::Integer = d.n)
    check(ccall((:vpm_sample_maxwellian, libvpm), Cint, (Ptr{Cvoid}, Int64, Int64, UInt64, Float64, Float64, Float64, Cint, Float64),
                d.h, offset, ntotal, seed, p.domain[1], p.domain[2], p.shift, 0, 1.0))
    d
end

# projection!(init::SplineDistribution, final::ParticleDistribution): an empty TODO upstream
# (src/projections/distribution.jl:57-61) -- stratified inverse-CDF resampling of the velocities from the spline
function projection!(init::DeviceSplineDistribution, final::DeviceParticleDistribution; seed::UInt64 = 0x000000005EED0001,
                     offset::Integer = 0, ntotal::Integer = final.n, jitter::Bool = false)
    mass = Ref{Float64}(0.0)
    check(ccall((:vpm_resample_v, libvpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}, Int64, Int64, UInt64, Cint, Ref{Float64}),
                init.h, C_NULL, final.h, offset, ntotal, seed, jitter ? 1 : 0, mass))
    final
end

end # module
