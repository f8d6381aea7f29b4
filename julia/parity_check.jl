# parity_check.jl — the check that turns "parity unpinned" into a number, for a machine that has BOTH a Julia
# toolchain with VlasovMethods.jl's dependencies AND a B200 with libvpm_b200.so.
#
# NOT EXECUTED IN THIS REPOSITORY (no Julia in the build image, SURVEY F3).  It runs the reference's own Julia path and
# the device path (julia/VPMB200.jl over the C ABI) on IDENTICAL particle arrays and prints the normwise relative
# differences the north star bounds by 1e-12 (one step, fp64; scatter order differs), on the quantities that do not
# depend on unpinned conventions (SURVEY 8c): E at particles, f_s and f_s' at particles, pushed particles, right-hand
# sides, scalar moments.  Coefficient vectors are compared up to the conventions the header documents (cyclic index
# shift of the periodic basis; gauge of the potential).
#
#   LIBVPM_B200=/path/to/libvpm_b200.so julia --project=/path/to/VlasovMethods.jl julia/parity_check.jl
#
# Expected output: every line "ok" (< 1e-12, or the stated looser bound for cond-amplified coefficient vectors).
using LinearAlgebra
using Random
using BSplineKit
using PoissonSolvers
using VlasovMethods

include(joinpath(@__DIR__, "VPMB200.jl"))
using .VPMB200

relerr(a, b) = norm(a .- b) / max(norm(b), floatmin())
function report(name, got, want; tol = 1e-12)
    e = relerr(got, want)
    println(rpad(name, 58), e < tol ? "ok   " : "FAIL ", e)
    e < tol
end
# best agreement over cyclic shifts (the periodic basis' index origin is a BSplineKit convention)
function relerr_cyclic(a, b)
    best, shift = Inf, 0
    for s in 0:length(a)-1
        e = relerr(circshift(a, s), b)
        e < best && ((best, shift) = (e, s))
    end
    best, shift
end

Random.seed!(20261017)
allok = true

# ------------------------------------------------------------------ Vlasov-Poisson: scripts/vlasov_poisson.jl:6-30
npart, nknot, order, domain, tstep = 10_000, 16, 3, (0.0, 1.0), 0.1
dist = initialize!(ParticleDistribution(1, 1, npart), NormalDistribution())
potential = Potential(PeriodicBasisBSplineKit(domain, order, nknot))
x = collect(dist.particles.x[1, :]); v = collect(dist.particles.v[1, :]); w = collect(dist.particles.w[1, :])

projection!(potential, dist)                       # src/projections/potential.jl:2-22
PoissonSolvers.update!(potential)                  # call site src/models/vlasov_poisson.jl:14
dphi_ref = [potential(xi, Derivative(1)) for xi in x]

# NOTE: whether PeriodicBasisBSplineKit(domain, order, nknot) has nknot or nknot-1 functions is unpinned (SURVEY 8c);
# take the size from the reference object.
nbasis = length(potential.rhs)
ddist = VPMB200.DeviceParticleDistribution(dist)
dpot = VPMB200.DevicePotential(domain, order, nbasis)
projection!(dpot, ddist)
VPMB200.update!(dpot)

e, s = relerr_cyclic(dpot.rhs, collect(potential.rhs))
println(rpad("deposit rhs (best cyclic shift = $s)", 58), e < 1e-12 ? "ok   " : "FAIL ", e); allok &= e < 1e-12
c_ref = collect(potential.coefficients); c_dev = circshift(dpot.coefficients, s)
allok &= report("potential coefficients (means removed)", c_dev .- sum(c_dev) / nbasis, c_ref .- sum(c_ref) / nbasis; tol = 1e-10)
allok &= report("phi'(x_p) at the particles", dpot(x, Derivative(1)), dphi_ref)

# one Strang step with the shipped (frozen-field) flows: drift/2, kick/2, kick/2, drift/2 (src/models/vlasov_poisson.jl:53-67,85)
model = VlasovPoisson(dist, potential)
params = (ϕ = potential, model = model)
z0 = copy(dist.particles.z); z1 = similar(z0); z2 = similar(z0)
VlasovMethods.s_advection!(z1, tstep / 2, z0, 0.0, params)
VlasovMethods.s_acceleration!(z2, tstep / 2, z1, 0.0, params)
VlasovMethods.s_acceleration!(z1, tstep / 2, z2, 0.0, params)
VlasovMethods.s_advection!(z2, tstep / 2, z1, 0.0, params)
dmodel = VlasovPoisson(ddist, dpot)
m = SplittingMethod(dmodel, (0.0, tstep), tstep)           # field = :frozen
run!(m)
zd = Matrix{Float64}(undef, 2, npart); VPMB200.download!(zd, ddist)
allok &= report("x after one Strang step (frozen field, as shipped)", zd[1, :], z2[1, :])
allok &= report("v after one Strang step (frozen field, as shipped)", zd[2, :], z2[2, :])

# ------------------------------------------------------------------ Lenard-Bernstein: scripts/lenard_bernstein_conservative.jl:10-27
nknotv, orderv, domainv, ν = 41, 4, (-10.0, 10.0), 1.0
ldist = initialize!(ParticleDistribution(1, 1, npart), DoubleMaxwellian(domainv, 2.0))
sdist = SplineDistribution(1, 1, nknotv, orderv, domainv, :Dirichlet)
vv = collect(ldist.particles.v[1, :])
fs = projection(vv, ldist, sdist)                   # src/projections/distribution.jl:35-55
dfs = Derivative(1) * fs

dl = VPMB200.DeviceParticleDistribution(ldist)
ds = VPMB200.DeviceSplineDistribution(nknotv, orderv, domainv, :Dirichlet)
dfsp = projection(vv, dl, ds)
allok &= report("projected coefficients (cond(M) = 20)", VPMB200.coefficients(ds), collect(sdist.coefficients); tol = 1e-11)
allok &= report("f_s(v_p)", dfsp(vv), fs.(vv))
allok &= report("f_s'(v_p)", (Derivative(1) * dfsp)(vv), dfs.(vv))
mref = vcat(collect(VlasovMethods.compute_f_densities(sdist, vv)), collect(VlasovMethods.compute_df_densities(sdist, vv)))
mdev = VPMB200.moments(ds, vv)
scale = [sum(abs, fs.(vv)), sum(abs, vv .* fs.(vv)), sum(abs, vv .^ 2 .* fs.(vv)), sum(abs, dfs.(vv)), sum(abs, vv .* dfs.(vv))]
allok &= report("five moments (relative to Σ|.|, SURVEY 8a note)", mdev ./ scale, mref ./ scale)

for (name, Model, rhs!) in (("LB_rhs!", LenardBernstein, VlasovMethods.LB_rhs!), ("CLB_rhs!", ConservativeLenardBernstein, VlasovMethods.CLB_rhs!))
    mref_ = Model(ldist, CollisionEntropy(sdist))
    pref = (ν = mref_.ν, idist = ldist, fdist = sdist, model = mref_)
    vdot_ref = similar(vv); rhs!(vdot_ref, vv, pref, 0.0)
    mdev_ = Model(dl, CollisionEntropy(ds))
    pdev = (ν = mdev_.ν, idist = dl, fdist = ds, model = mdev_)
    vdot_dev = similar(vv); rhs!(vdot_dev, vv, pdev, 0.0)
    global allok &= report(name, vdot_dev, vdot_ref; tol = 1e-10)
    # five RK438 steps through the reference's GeometricIntegrator against the device stepper
    gi = GeometricIntegrator(mref_, (0.0, 5e-2), 1e-2)
    tmp = tempname() * ".hdf5"; run!(gi, tmp)
    v_ref = collect(ldist.particles.v[1, :])
    VPMB200.upload!(dl, Matrix(vcat(zeros(1, npart), vv', fill(1 / npart, 1, npart))))
    run!(GeometricIntegrator(mdev_, (0.0, 5e-2), 1e-2))
    zl = Matrix{Float64}(undef, 3, npart); VPMB200.download!(zl, dl)
    global allok &= report("$(name[1:end-5]) RK438, 5 steps", zl[2, :], v_ref; tol = 1e-11)
    ldist.particles.v[1, :] .= vv                   # restore the reference's particles for the next model
end

println(allok ? "\nPARITY PINNED: all quantities within tolerance" : "\nPARITY CHECK FAILED: see the FAIL lines")
exit(allok ? 0 : 1)
