#!/bin/bash
# round-2 first probe: sanity, sustained windows of the three steppers, box topology
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
(lscpu; numactl --hardware 2>&1; nproc; free -g) > gpurun_out/r2_host.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_probe_tests.log
python bench.py --steps 100 --warmup 5 > gpurun_out/r2_probe_vp_100.json 2> gpurun_out/r2_probe_vp_100.err
python bench.py --steps 8000 --warmup 5 --no-e2e --no-cpu > gpurun_out/r2_probe_vp_sustained.json 2>> gpurun_out/r2_probe_vp_100.err
python bench.py --workload lb --steps 50 --no-cpu > gpurun_out/r2_probe_lb_50.json 2>> gpurun_out/r2_probe_vp_100.err
python bench.py --workload lb --steps 2500 --no-cpu > gpurun_out/r2_probe_lb_sustained.json 2>> gpurun_out/r2_probe_vp_100.err
python bench.py --workload clb --steps 50 --no-cpu > gpurun_out/r2_probe_clb_50.json 2>> gpurun_out/r2_probe_vp_100.err
python bench.py --workload clb --steps 1800 --no-cpu > gpurun_out/r2_probe_clb_sustained.json 2>> gpurun_out/r2_probe_vp_100.err
tail -3 gpurun_out/r2_probe_tests.log
cat gpurun_out/r2_probe_vp_sustained.json
