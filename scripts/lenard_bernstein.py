#!/usr/bin/env python
"""scripts/lenard_bernstein.jl on the GPU (BASELINE config 3): collisional relaxation with the spline projection of the
entropy gradient, nknot 41, order 4, v in (-10, 10), NormalDistribution initial data, t in (0, 10), dt 0.1.  The shipped
script integrates with DiffEqIntegrator/TRBDF2 through ForwardDiff duals, which cannot scale past ~1e3 particles
(SURVEY F6); this mirror uses the package's other integrator for these models, GeometricIntegrator (RK438).  Like the
shipped script (:27-28) it instantiates ConservativeLenardBernstein; --plain selects LenardBernstein.

    python scripts/lenard_bernstein.py [--npart 1000] [--plain]
"""
import argparse

import numpy as np

import _common  # noqa: F401
from vpm_b200 import (CollisionEntropy, ConservativeLenardBernstein, GeometricIntegrator, LenardBernstein, NormalDistribution,
                      ParticleDistribution, SplineDistribution, initialize_, run_)

ap = argparse.ArgumentParser()
ap.add_argument("--npart", type=float, default=1000)
ap.add_argument("--plain", action="store_true")
args = ap.parse_args()

# parameters                                                          scripts/lenard_bernstein.jl:11-16
npart = int(args.npart)   # number of particles
nknot = 41                # number of grid points
order = 4                 # spline order
tstep = 0.1               # time step size
tspan = (0.0, 1e1)        # integration time interval
domainv = (-10.0, 10.0)

# create and initialize particle distribution function                :19
dist = initialize_(ParticleDistribution(1, 1, npart), NormalDistribution())

# create spline distribution function and entropy                     :22-23
sdist = SplineDistribution(1, 1, nknot, order, domainv, "Dirichlet")
entropy = CollisionEntropy(sdist)

# create LenardBernstein model                                        :26-28
model = (LenardBernstein if args.plain else ConservativeLenardBernstein)(dist, entropy)

# create integrator                                                   :31
integrator = GeometricIntegrator(model, tspan, tstep)

print("Running integrator")
run_(integrator)                                                      # :38

sv, sv2 = integrator.diagnostics.T
for n in range(0, len(sv), 10):
    print(f"step {n:3d}: sum v = {sv[n]:+.9e}  sum v^2 = {sv2[n]:.9e}")
print(f"momentum change / N: {abs(sv[-1] - sv[0]) / npart:.2e}; energy change: {abs(sv2[-1] - sv2[0]) / sv2[0]:.2e}")
