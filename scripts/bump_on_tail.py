#!/usr/bin/env python
"""scripts/bump_on_tail.jl on the GPU: bump-on-tail instability (BASELINE config 2; 1e8 particles fit one B200).
Parameters of scripts/bump_on_tail.jl:14-30; the legacy self-consistent loop integrate_vp! (src/vlasov_poisson.jl:94-115)
is the "selfconsistent" mode of the Strang stepper.  Prints W, K, W + K (the script's three log plots, :64-71) and the
fitted growth rate of the field energy.

    python scripts/bump_on_tail.py [--npart 50000] [--T 50]
"""
import argparse

import numpy as np

import _common  # noqa: F401
from vpm_b200 import (BumpOnTail, tspan_for, ParticleDistribution, PeriodicBasisBSplineKit, Potential, SplittingMethod, VlasovPoisson,
                      initialize_, run_)

ap = argparse.ArgumentParser()
ap.add_argument("--npart", type=float, default=5e4)
ap.add_argument("--T", type=float, default=50.0)
args = ap.parse_args()

# simulation parameters                                              scripts/bump_on_tail.jl:14-30
dt = 1e-1                   # timestep
T = args.T                  # final time
nt = int(T // dt)           # nb. of timesteps: Int(div(T, dt)) as upstream (= 499 for T = 50, since 0.1 > 1/10)
nh = 16                     # nb. of elements
p = 3                       # spline degree (order p + 1)
npart = int(args.npart)     # nb. of particles
params = dict(kappa=0.3, eps=0.03, alpha=0.1, v0=4.5, sigma=0.5)
chi = 1.0
L = 2 * np.pi / params["kappa"]   # domain length

# initial data (device-side, counter-based; bumpontail.jl:43-75)       :44
dist = initialize_(ParticleDistribution(1, 1, npart), BumpOnTail(**params))

# B-spline Poisson solver + scaled field                              :40-41
potential = Potential(PeriodicBasisBSplineKit((0.0, L), p + 1, nh))
model = VlasovPoisson(dist, potential)

# integrate all time steps                                            :58-60
integrator = SplittingMethod(model, tspan_for(nt, dt), dt, field="selfconsistent", chi=chi)
run_(integrator, diag_mode=2)          # W, K, M exactly as save_timestep! (src/vlasov_poisson.jl:58-67)

W, K, M = integrator.diagnostics.T
t = dt * np.arange(nt + 1)
for n in range(0, nt + 1, max(nt // 10, 1)):
    print(f"t = {t[n]:5.1f}  W = {W[n]:.6e}  K = {K[n]:.6e}  W+K = {W[n] + K[n]:.9e}  M = {M[n]:+.6e}")
print(f"W(0) = {W[0]:.6e}, max W = {W.max():.6e} at t = {t[W.argmax()]:.1f}")
print(f"relative energy drift: {abs((W + K)[-1] - (W + K)[0]) / (W + K)[0]:.2e}")
# growth of the field energy on the linear phase (between 3x the initial level and 1/20 of the maximum)
lo, hi = np.argmax(W > 3 * W[:10].min()), np.argmax(W > W.max() / 20)
if hi > lo + 10:
    gamma = 0.5 * np.polyfit(t[lo:hi], np.log(W[lo:hi]), 1)[0]
    print(f"fitted growth rate gamma = {gamma:.4f} on t in ({t[lo]:.1f}, {t[hi]:.1f})  (dispersion relation: 0.198)")
