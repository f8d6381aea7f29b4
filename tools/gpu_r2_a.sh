#!/bin/bash
# round 2, run A: all GPU tests with the parity-error recorder, the new default bench line, the reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2_tests_a.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2_ref_a.json 2> gpurun_out/r2_ref_a.err
tail -15 gpurun_out/r2_tests_a.log; tail -5 gpurun_out/r2_bench_a.err; tail -4 gpurun_out/r2_ref_a.err
