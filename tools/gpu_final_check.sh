# what the driver does at round end, on one GPU: build artefacts are in-tree; tests, smoke, default bench, reference arm
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/gpu_tests.log
python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/smoke.log
if [ "$1" != "quick" ]; then
python bench.py 2>&1 | tail -1 > gpurun_out/final_bench_default.json; cut -c1-250 gpurun_out/final_bench_default.json
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-200
fi
