#!/bin/bash
# round 2, run D: tests (incl. entropy), default bench line, LB stage-release A/B
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_tests_d.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err
for w in lb clb; do
VPM_TUNE_LBREL=1 python bench.py --workload $w --steps 30 --no-cpu --no-extras > gpurun_out/r2_ab_${w}_late.json 2>> gpurun_out/r2_bench_d.err
done
tail -8 gpurun_out/r2_tests_d.log; tail -5 gpurun_out/r2_bench_d.err
