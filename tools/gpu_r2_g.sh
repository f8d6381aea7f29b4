#!/bin/bash
# round 2, run G: ncu --set full of one CLB step's passes (thin ring), to see what actually bounds the moments pass and stage 1
mkdir -p gpurun_out
VPM_TUNE_LBFAT=0 ncu --set full --clock-control none --import-source on -k regex:"lb_pass" -s 9 -c 8 \
    -o gpurun_out/r2_ncu_clb_passes python bench.py --workload clb --steps 4 --warmup 3 --no-cpu --no-extras > gpurun_out/r2_ncu_clb.log 2>&1
ls -la gpurun_out/r2_ncu_clb_passes.ncu-rep
python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -5
