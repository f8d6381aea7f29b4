#!/bin/bash
# round-2 final evidence in one call: default bench line (the driver's command), launch lists of the same commands,
# ncu --set full of the dominant kernels (VP ring pass incl. the carried last pass; velocity-sorted LB passes + field kernel), smoke
set -x
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/r2r.err | tail -1 > gpurun_out/r2r_bench_default_n1.json; cut -c1-300 gpurun_out/r2r_bench_default_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2r_launches_vp.csv python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_launch_vp.log 2>&1
for w in lb clb; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2r_launches_${w}.csv python bench.py --workload $w --steps 4 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_launch_${w}.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:vp_pass -s 6 -c 4 -o gpurun_out/r2r_prof_vp python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_full_vp.log 2>&1
python __graft_entry__.py smoke 2>&1 | tail -1
du -sh gpurun_out
