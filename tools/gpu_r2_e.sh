#!/bin/bash
# round 2, run E: tests (entropy, tiled v-deposit), LB/CLB after the fp64 diet, large v-grid: tiled deposit vs CAS fallbacks
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_tests_e.log
for w in lb clb; do
python bench.py --workload $w --steps 30 --no-cpu --no-extras > gpurun_out/r2_e_${w}.json 2>> gpurun_out/r2_e.err
done
python bench.py --workload lb --steps 10 --no-cpu --no-extras --lb-nknots 300 > gpurun_out/r2_e_lb_nk300_tiled.json 2>> gpurun_out/r2_e.err
VPM_TUNE_HM=1 python bench.py --workload lb --steps 10 --no-cpu --no-extras --lb-nknots 300 > gpurun_out/r2_e_lb_nk300_cas.json 2>> gpurun_out/r2_e.err
python bench.py --workload lb --steps 10 --no-cpu --no-extras --lb-nknots 1000 > gpurun_out/r2_e_lb_nk1000_tiled.json 2>> gpurun_out/r2_e.err
tail -8 gpurun_out/r2_tests_e.log; tail -5 gpurun_out/r2_e.err
