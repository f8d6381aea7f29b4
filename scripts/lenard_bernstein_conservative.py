#!/usr/bin/env python
"""scripts/lenard_bernstein_conservative.jl on the GPU: conservative Lenard-Bernstein relaxation of a double Maxwellian
(BASELINE config 4).  Line for line scripts/lenard_bernstein_conservative.jl:1-64 up to the animation; the per-frame
momentum / energy prints of the animation loop (:49-50,64) are kept.

    python scripts/lenard_bernstein_conservative.py [--npart 1000] [--tend 500] [--plain]
"""
import argparse

import numpy as np

from _common import h5read
from vpm_b200 import (CollisionEntropy, ConservativeLenardBernstein, DoubleMaxwellian, GeometricIntegrator, LenardBernstein,
                      ParticleDistribution, SplineDistribution, initialize_, projection, run_)

ap = argparse.ArgumentParser()
ap.add_argument("--npart", type=float, default=1000)
ap.add_argument("--tend", type=float, default=5e2)
ap.add_argument("--save-stride", type=int, default=500)
ap.add_argument("--plain", action="store_true", help="LenardBernstein instead of ConservativeLenardBernstein (:28)")
ap.add_argument("--h5file", default="lenard_bernstein_conservative.hdf5")
args = ap.parse_args()

# output file                                                         scripts/lenard_bernstein_conservative.jl:6
h5file = args.h5file

# parameters                                                          :10-15
npart = int(args.npart)    # number of particles
nknot = 41                 # number of grid points
order = 4                  # spline order
tstep = 1e-2               # time step size
tspan = (0.0, args.tend)   # integration time interval
domainv = (-10.0, 10.0)

# create and initialize particle distribution function                :18
dist = initialize_(ParticleDistribution(1, 1, npart), DoubleMaxwellian(domainv, 2.0))

# create spline distribution function and entropy                     :22-23
sdist = SplineDistribution(1, 1, nknot, order, domainv, "Dirichlet")
entropy = CollisionEntropy(sdist)

# create LenardBernstein model                                        :26-27
model = (LenardBernstein if args.plain else ConservativeLenardBernstein)(dist, entropy)

# create integrator (RK438, as shipped)                               :30
integrator = GeometricIntegrator(model, tspan, tstep)

print("Running integrator")
run_(integrator, h5file, save_stride=args.save_stride)                # :39

# read array from HDF5 file                                           :46-47
z = h5read(h5file, "z")
t = h5read(h5file, "t")

mom = z.sum(axis=0)                                                   # :49
enr = (z ** 2).sum(axis=0)                                            # :50
for n in range(z.shape[1]):
    print(f"t = {t[n]:7.2f}, mom = {(mom[n] - mom[0]) / mom[0]:+.3e}, enr = {(enr[n] - enr[0]) / enr[0]:+.3e}, "
          f"min(v) = {z[:, n].min():.3f}, max(v) = {z[:, n].max():.3f}")   # :64
# the projected distribution of the last frame (:67) and its kurtosis: 1.72 for the double Maxwellian, 3 at equilibrium
f = projection(z[:, -1], dist, sdist)
vgrid = np.linspace(-8, 8, 9)
print("f_s on", vgrid, "=", np.array2string(f(vgrid), precision=4))
c = z[:, -1] - z[:, -1].mean()
print(f"normalised fourth moment: {(c ** 4).mean() / (c ** 2).mean() ** 2:.4f}")
