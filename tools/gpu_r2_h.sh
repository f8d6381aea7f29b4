#!/bin/bash
# round 2, run H: tests, CLB/LB with cp.async gather passes, e2e chunking A/B, compute-sanitizer on every pass variant
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_tests_h.log
for w in lb clb; do
python bench.py --workload $w --steps 30 --no-cpu --no-extras --sustained > gpurun_out/r2_h_${w}.json 2>> gpurun_out/r2_h.err
done
VPM_TUNE_E2E_CHUNKS=1 python bench.py --steps 20 --no-cpu --no-extras --e2e-steps 10 > gpurun_out/r2_h_e2e_chunks1.json 2>> gpurun_out/r2_h.err
VPM_TUNE_E2E_CHUNKS=8 python bench.py --steps 20 --no-cpu --no-extras --e2e-steps 10 > gpurun_out/r2_h_e2e_chunks8.json 2>> gpurun_out/r2_h.err
python tests/sanitize_workload.py 2>&1 | tail -2
for tool in memcheck racecheck synccheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tests/sanitize_workload.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload|Error|error" gpurun_out/r2_sanitizer_$tool.log | head -6
done
tail -8 gpurun_out/r2_tests_h.log; tail -5 gpurun_out/r2_h.err
