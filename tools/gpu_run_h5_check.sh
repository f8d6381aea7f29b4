# run! drivers with HDF5 trajectory output: GPU tests, the overlap measurement (tmpfs and disk), smoke, whole GPU suite
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_run_h5.py -x -q 2>&1 | tail -15 | tee gpurun_out/h5_tests.log
timeout 200 python tools/run_h5_overlap.py --particles 20000000 2>&1 | tail -1 | tee gpurun_out/r1_run_h5_overlap_shm.json
timeout 200 python tools/run_h5_overlap.py --particles 20000000 --dir /tmp 2>&1 | tail -1 | tee gpurun_out/r1_run_h5_overlap_tmp.json
python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/smoke.log
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_run_h5.py 2>&1 | tail -3 | tee gpurun_out/gpu_tests.log
