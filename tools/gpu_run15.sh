set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for pair in 1 0; do
for wl in lb clb; do
  VPM_TUNE_LBPAIR=$pair python bench.py --workload $wl --steps 20 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); r=d['roofline']; print('PAIR', $pair, '$wl', 'ms/step %.4f'%d['ms_per_step'], 'GB/s %.0f'%r['achieved'], 'frac %.3f'%r['frac'])"
done; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_clb_e.csv python bench.py --workload clb --steps 2 --warmup 3 > gpurun_out/ncu_launch_lb.log 2>&1
