set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -2
for t in -1 0; do
for w in lb clb; do
VPM_TUNE_LBTMA=$t python bench.py --workload $w --steps 50 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/bench_${w}_i$t.json
python - <<P
import json
d=json.load(open("gpurun_out/bench_${w}_i$t.json")); r=d["roofline"]
print("RESULT", "$w", "tma=$t", "ms/step %.4f"%d["ms_per_step"], "%.4e"%d["value"], "pass ms %.4f"%r["avg_launch_ms"])
P
done
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_clb_i.csv python bench.py --workload clb --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch_clb.log 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_clb_i.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows[-16:]:
    if 'field' not in r[4]: print(r[4][:60], r[-1])
P
