#!/bin/bash
# round-2 evidence in one call: default bench line (the driver's command), reference arm, launch lists of the same commands,
# ncu --set full of the dominant kernels, smoke
set -x
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/r2c.err | tail -1 > gpurun_out/r2c_bench_default_n1.json; cut -c1-300 gpurun_out/r2c_bench_default_n1.json
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>>gpurun_out/r2c.err | tail -1 > gpurun_out/r2c_bench_reference_n1.json; cut -c1-300 gpurun_out/r2c_bench_reference_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2c_launches_vp.csv python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_launch_vp.log 2>&1
for w in lb clb; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2c_launches_${w}.csv python bench.py --workload $w --steps 4 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_launch_${w}.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:vp_ -s 8 -c 2 -o gpurun_out/r2c_prof_vp python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_full_vp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lb_ -s 17 -c 16 -o gpurun_out/r2c_prof_clb python bench.py --workload clb --steps 4 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_full_clb.log 2>&1
python __graft_entry__.py smoke 2>&1 | tail -1
du -sh gpurun_out
