set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 100 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/bench_vp_b.json; cat gpurun_out/bench_vp_b.json
python bench.py --workload lb --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_lb_b.json; cat gpurun_out/bench_lb_b.json
python bench.py --workload clb --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_clb_b.json; cat gpurun_out/bench_clb_b.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_lb_b.csv python bench.py --workload clb --steps 2 --warmup 3 > gpurun_out/ncu_launch_lb.log 2>&1
