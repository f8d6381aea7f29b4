#!/bin/bash
# round 2, run C: why is the instruction-light ring pass slower at burst clocks?  A/B of the stage-release policy + ncu of both passes
mkdir -p gpurun_out
B="python bench.py --steps 100 --no-e2e --no-cpu --no-extras"
VPM_TUNE_TMA=5 VPM_TUNE_VPREL=1 $B --sustained > gpurun_out/r2_ab_tma5_late.json 2> gpurun_out/r2_c.err
VPM_TUNE_TMA=5 VPM_TUNE_VPREL=0 $B > gpurun_out/r2_ab_tma5_early.json 2>> gpurun_out/r2_c.err
VPM_TUNE_TMA=1 $B > gpurun_out/r2_ab_tma1_again.json 2>> gpurun_out/r2_c.err
for v in 1 5; do
  VPM_TUNE_TMA=$v ncu --set full --clock-control none --import-source on -k regex:"vp_pass_(tma|ring)_kernel" -s 4 -c 1 \
      -o gpurun_out/r2_ncu_vp_tma$v python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r2_ncu_tma$v.log 2>&1
done
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/r2_c.err
