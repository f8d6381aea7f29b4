#!/usr/bin/env python
"""scripts/vlasov_poisson.jl on the GPU: 1D1V Vlasov-Poisson PIC with a spline Poisson solve, as shipped
(BASELINE config 1).  Line for line the reference script (scripts/vlasov_poisson.jl:1-38) with Julia's `f!` spelled
`f_`; only the import changes.  The plots of the original (:40-75) are replaced by printed diagnostics.

    python scripts/vlasov_poisson.py [--npart 10000] [--field frozen|selfconsistent] [--save-stride 1]
"""
import argparse

import numpy as np

from _common import h5read
from vpm_b200 import (NormalDistribution, ParticleDistribution, PeriodicBasisBSplineKit, Potential, SplittingMethod,
                      VlasovPoisson, initialize_, run_)

ap = argparse.ArgumentParser()
ap.add_argument("--npart", type=int, default=10000)
ap.add_argument("--field", default="frozen", help="frozen = the shipped behaviour (SURVEY F4); selfconsistent = physical loop")
ap.add_argument("--save-stride", type=int, default=1, help="1 = every step, as upstream")
ap.add_argument("--h5file", default="vlasov_poisson.hdf5")
args = ap.parse_args()

# parameters                                                        scripts/vlasov_poisson.jl:6-11
npart = args.npart     # number of particles
nknot = 16             # number of grid points
order = 3              # spline order
tstep = 0.1            # time step size
tspan = (0.0, 20.0)    # integration time interval
domain = (0.0, 1.0)

# output file                                                       :14
h5file = args.h5file

# create and initialize particle distribution function               :17
dist = initialize_(ParticleDistribution(1, 1, npart), NormalDistribution())

# create electrostatic potential                                     :21
potential = Potential(PeriodicBasisBSplineKit(domain, order, nknot))

# create Vlasov-Poisson model                                        :24
model = VlasovPoisson(dist, potential)

# create integrator                                                  :27
integrator = SplittingMethod(model, tspan, tstep, field=args.field)

# integrate                                                          :30
run_(integrator, h5file, save_stride=args.save_stride)

# read array from HDF5 file                                          :38
z = h5read(h5file, "z")

# compute plot ranges                                                :41-43
vmax = np.ceil(np.abs(z[1, :, 0]).max())
print(f"z: {z.shape} (nd, np, frames); vmax = {vmax}")
W, K, M = integrator.diagnostics.T
for n in range(0, len(K), max(len(K) // 10, 1)):
    print(f"step {n:4d}  W = {W[n]:.6e}  K = {K[n]:.6e}  M = {M[n]:+.6e}")
x_end, v_end, _ = dist.get()
assert np.array_equal(z[0, :, -1], x_end) and np.array_equal(z[1, :, -1], v_end)   # last frame = final state (:49)
print(f"mean x mod 1 = {np.mod(z[0, :, -1], 1).mean():.4f}, rms v = {z[1, :, -1].std():.4f}")
