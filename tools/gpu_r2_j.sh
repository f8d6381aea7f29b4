#!/bin/bash
# round 2, run J: tests; config 4 (CLB) and plain LB with the entropy history: 1e7 particles x 5e4 steps, and 1e8 x 2000 steps
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2_tests_j.log
python tools/lb_checks.py 1e7 50000 clb,lb entropy > gpurun_out/r2_physics_entropy_1e7.json 2> gpurun_out/r2_j.err
python tools/lb_checks.py 1e8 2000 clb entropy > gpurun_out/r2_physics_entropy_1e8.json 2>> gpurun_out/r2_j.err
tail -4 gpurun_out/r2_tests_j.log; tail -3 gpurun_out/r2_j.err; cut -c1-600 gpurun_out/r2_physics_entropy_1e8.json
