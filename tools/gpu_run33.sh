set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in lb clb; do
python bench.py --workload $w --steps 30 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/bench_${w}_v5.json; python -c "
import json; d=json.load(open('gpurun_out/bench_${w}_v5.json')); print('RESULT $w ms/step %.4f  %.3e p-steps/s  pass GB/s %.0f'%(d['ms_per_step'], d['value'], d['roofline']['achieved']))"
done
ncu --set full --clock-control none --import-source on -k regex:lb_pass -s 5 -c 4 -o gpurun_out/prof_lb_v5 python bench.py --workload lb --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_lb_v5.log 2>&1
