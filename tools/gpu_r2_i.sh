#!/bin/bash
# round 2, run I: tests; VP without diagnostics accumulation (burst + sustained); CLB with 4-particle gather trips vs 2
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_tests_i.log
python bench.py --steps 100 --no-e2e --no-cpu --no-extras --sustained > gpurun_out/r2_i_vp.json 2> gpurun_out/r2_i.err
python bench.py --workload clb --steps 30 --no-cpu --no-extras > gpurun_out/r2_i_clb_np4.json 2>> gpurun_out/r2_i.err
VPM_TUNE_LBNP=2 python bench.py --workload clb --steps 30 --no-cpu --no-extras > gpurun_out/r2_i_clb_np2.json 2>> gpurun_out/r2_i.err
tail -8 gpurun_out/r2_tests_i.log; tail -5 gpurun_out/r2_i.err
