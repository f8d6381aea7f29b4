#!/bin/bash
# round 2, run L: the BASELINE configs at full size with the round-2 kernels: config 2 (1e8 particles: Landau / bump-on-tail rates),
# config 4 (CLB, 1e8 particles x 5e4 RK438 steps: conservation over the whole run, sustained throughput)
mkdir -p gpurun_out
( time python tools/physics_checks.py 1e8 ) > gpurun_out/r2_physics_1e8.json 2> gpurun_out/r2_l.err
( time python tools/lb_checks.py 1e8 50000 clb ) > gpurun_out/r2_physics_clb_1e8.json 2>> gpurun_out/r2_l.err
tail -12 gpurun_out/r2_l.err; cut -c1-400 gpurun_out/r2_physics_1e8.json
