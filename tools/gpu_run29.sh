set -x
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_uniform_weight.py::test_lb_uniform_weight 2>&1 | tail -3
for t in -1 0; do
for w in lb clb; do
VPM_TUNE_LBTMA=$t python bench.py --workload $w --steps 30 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/bench_${w}_t$t.json; python -c "
import json; d=json.load(open('gpurun_out/bench_${w}_t$t.json')); print('RESULT $w tma=$t ms/step %.4f  %.3e p-steps/s  pass GB/s %.0f'%(d['ms_per_step'], d['value'], d['roofline']['achieved']))"
done
VPM_TUNE_LBTMA=$t ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_clb_t$t.csv python bench.py --workload clb --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch_clb.log 2>&1
python - <<P
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_clb_t$t.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows[-16:]:
    if 'field' not in r[4]: print("tma=$t", r[4][:70], r[-1])
P
done
