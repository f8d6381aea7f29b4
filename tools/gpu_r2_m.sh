#!/bin/bash
# round 2, run M: velocity-sorted collision passes -- ring depth / release policy sweep, then BASELINE config 4 at full size and
# length (CLB, 1e8 particles x 5e4 RK438 steps) on the sorted path: conservation over the whole run, sustained throughput
mkdir -p gpurun_out
for ring in 16 32 48 64 80; do for rel in 0 1; do
  VPM_TUNE_LBSRING=$ring VPM_TUNE_LBSREL=$rel timeout 200 python bench.py --workload clb --steps 30 --no-cpu --no-e2e --no-extras > gpurun_out/r2m_sweep_${ring}_${rel}.json 2>/dev/null
done; done
( time python tools/lb_checks.py 1e8 50000 clb ) > gpurun_out/r2m_physics_clb_1e8_sorted.json 2> gpurun_out/r2m.err
tail -5 gpurun_out/r2m.err
