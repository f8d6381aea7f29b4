set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/r1f_bench_vp_n1.json; cut -c1-300 gpurun_out/r1f_bench_vp_n1.json
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/r1f_bench_reference_n1.json
for w in lb clb; do
python bench.py --workload $w --steps 50 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/r1f_bench_${w}_n1.json; cut -c1-200 gpurun_out/r1f_bench_${w}_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1f_launches_${w}.csv python bench.py --workload $w --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch_${w}.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1f_launches_vp.csv python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch_vp.log 2>&1
ncu --set full --clock-control none -k regex:lb_ -s 17 -c 4 -o gpurun_out/r1f_prof_clb python bench.py --workload clb --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_clb.log 2>&1
ncu --set full --clock-control none -k regex:lb_pass -s 5 -c 4 -o gpurun_out/r1f_prof_lb python bench.py --workload lb --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_lb.log 2>&1
ncu --set full --clock-control none -k regex:vp_ -s 8 -c 2 -o gpurun_out/r1f_prof_vp python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_vp.log 2>&1
python __graft_entry__.py smoke 2>&1 | tail -1
du -sh gpurun_out
