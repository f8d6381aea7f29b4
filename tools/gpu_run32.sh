set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for t in -1 46 174; do
for w in lb clb; do
VPM_TUNE_LBTMA=$t python bench.py --workload $w --steps 30 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/bench_${w}_t$t.json; python -c "
import json; d=json.load(open('gpurun_out/bench_${w}_t$t.json')); print('RESULT $w tma=$t ms/step %.4f  %.3e p-steps/s  pass GB/s %.0f'%(d['ms_per_step'], d['value'], d['roofline']['achieved']))"
done
done
VPM_TUNE_LBTMA=190 ncu --set full --clock-control none --import-source on -k regex:lb_pass -s 9 -c 8 -o gpurun_out/prof_clb_v4 python bench.py --workload clb --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_clb_v4.log 2>&1
VPM_TUNE_LBTMA=0 ncu --set full --clock-control none --import-source on -k regex:lb_pass -s 9 -c 8 -o gpurun_out/prof_clb_v4_reg python bench.py --workload clb --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_clb_v4r.log 2>&1
