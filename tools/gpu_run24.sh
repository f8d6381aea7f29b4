set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_vp_f.json; cat gpurun_out/bench_vp_f.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_vp_f.csv python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch_vp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vp_pass_tma_kernel -s 2 -c 1 -o gpurun_out/prof_vp_f python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_vp.log 2>&1
python __graft_entry__.py smoke 2>&1 | tail -2
