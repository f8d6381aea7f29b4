set -x
python tests/sanitize_workload.py 2>&1 | tail -2
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tests/sanitize_workload.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload|Error|error" gpurun_out/sanitizer_$tool.log | head -8
done
