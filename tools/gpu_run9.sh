set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_vp_c.json; cat gpurun_out/bench_vp_c.json
python bench.py --workload lb --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_lb_c.json; cat gpurun_out/bench_lb_c.json
python bench.py --workload clb --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_clb_c.json; cat gpurun_out/bench_clb_c.json
python bench.py --impl reference --steps 10 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref_c.json; cat gpurun_out/bench_ref_c.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_vp_c.csv python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch_vp.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_clb_c.csv python bench.py --workload clb --steps 2 --warmup 3 > gpurun_out/ncu_launch_lb.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vp_pass_kernel -s 6 -c 1 -o gpurun_out/prof_vp_c python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_vp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lb_pass_kernel -s 12 -c 6 -o gpurun_out/prof_lb_c python bench.py --workload clb --steps 2 --warmup 3 > gpurun_out/ncu_full_lb.log 2>&1
ncu --set full --clock-control none -k regex:field_kernel -s 4 -c 2 -o gpurun_out/prof_field_c python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_field.log 2>&1
