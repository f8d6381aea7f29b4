set -x
python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_vp.json; cat gpurun_out/bench_vp.json
python bench.py --workload lb --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_lb.json; cat gpurun_out/bench_lb.json
python bench.py --workload clb --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_clb.json; cat gpurun_out/bench_clb.json
python tools/physics_checks.py 2e7 2>&1 | tail -2 > gpurun_out/physics.json; cat gpurun_out/physics.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_lb_r1.csv python bench.py --workload clb --steps 2 --warmup 3 > gpurun_out/ncu_launch_lb.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lb_pass_kernel -s 12 -c 5 -o gpurun_out/prof_lb_r1 python bench.py --workload clb --steps 2 --warmup 3 > gpurun_out/ncu_full_lb.log 2>&1
