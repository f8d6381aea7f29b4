#!/bin/bash
# round 2, run B: GPU tests on the warp-specialised VP ring pass, default bench line, A/B of the two TMA passes (burst + sustained)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_tests_b.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
VPM_TUNE_TMA=1 python bench.py --steps 100 --no-e2e --no-cpu --no-extras --sustained > gpurun_out/r2_ab_tma1.json 2>> gpurun_out/r2_bench_b.err
VPM_TUNE_TMA=5 python bench.py --steps 100 --no-e2e --no-cpu --no-extras --sustained > gpurun_out/r2_ab_tma5.json 2>> gpurun_out/r2_bench_b.err
VPM_TUNE_TMA=5 python bench.py --steps 100 --no-e2e --no-cpu --no-extras --sustained --n-basis 17 > gpurun_out/r2_ab_tma5_nh17.json 2>> gpurun_out/r2_bench_b.err
tail -8 gpurun_out/r2_tests_b.log; tail -5 gpurun_out/r2_bench_b.err
