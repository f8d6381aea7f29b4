# run! drivers with HDF5 trajectory output: overlap measurement (tmpfs and disk) and compute-sanitizer on the new paths
set -x
mkdir -p gpurun_out
timeout 200 python tools/run_h5_overlap.py --particles 20000000 2>&1 | tail -1 | tee gpurun_out/r1_run_h5_overlap_shm.json
timeout 200 python tools/run_h5_overlap.py --particles 20000000 --dir /tmp 2>&1 | tail -1 | tee gpurun_out/r1_run_h5_overlap_tmp.json
timeout 100 python tests/sanitize_workload.py --run-h5 2>&1 | tail -2
for tool in memcheck initcheck racecheck; do
  timeout 240 compute-sanitizer --tool $tool --print-limit 20 python tests/sanitize_workload.py --run-h5 > gpurun_out/sanitizer_h5_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload|Error|error" gpurun_out/sanitizer_h5_$tool.log | head -8
done
