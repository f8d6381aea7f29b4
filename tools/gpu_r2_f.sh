#!/bin/bash
# round 2, run F: tests (incl. the one-GPU two-rank p2p test), LB/CLB: fat ring CTA + replicated tables vs the 2-CTA ring, default line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_tests_f.log
for w in lb clb; do
python bench.py --workload $w --steps 30 --no-cpu --no-extras --sustained > gpurun_out/r2_f_${w}_fat.json 2>> gpurun_out/r2_f.err
VPM_TUNE_LBFAT=0 python bench.py --workload $w --steps 30 --no-cpu --no-extras > gpurun_out/r2_f_${w}_thin.json 2>> gpurun_out/r2_f.err
done
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_f_default.json 2>> gpurun_out/r2_f.err
tail -8 gpurun_out/r2_tests_f.log; tail -5 gpurun_out/r2_f.err
