# run! drivers with HDF5 trajectory output: new GPU tests, the overlap measurement, smoke, default bench, then the whole GPU suite
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_run_h5.py -x -q 2>&1 | tail -15 | tee gpurun_out/h5_tests.log
timeout 200 python tools/run_h5_overlap.py --particles 20000000 2>&1 | tail -1 | tee gpurun_out/r1_run_h5_overlap.json
python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/smoke.log
python bench.py 2>&1 | tail -1 > gpurun_out/final_bench_default.json; cut -c1-250 gpurun_out/final_bench_default.json
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_run_h5.py --durations=8 2>&1 | tail -14 | tee gpurun_out/gpu_tests.log
