set -x
for m in 3 2 4; do VPM_TUNE_MINB=$m python bench.py --steps 50 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/bench_minb$m.json; cat gpurun_out/bench_minb$m.json | python -c "import json,sys; d=json.load(sys.stdin); print('MINB',$m, d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['clocks'])"; done
python bench.py --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_full.json; cat gpurun_out/bench_full.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --particles 1e8 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vp_pass_kernel -s 6 -c 2 -o gpurun_out/prof_vp_r1 python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e --particles 1e8 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
