set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_samplers.py -m gpu -x -q 2>&1 | tail -2
python bench.py --workload clb --steps 50 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/bench_clb_j.json
python - <<P
import json
d=json.load(open("gpurun_out/bench_clb_j.json")); r=d["roofline"]
print("RESULT clb ms/step %.4f"%d["ms_per_step"], "%.4e"%d["value"], "pass ms %.4f"%r["avg_launch_ms"])
P
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_clb_j.csv python bench.py --workload clb --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch_clb.log 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_clb_j.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows[-8:]:
    if 'field' not in r[4]: print(r[4][:60], r[-1])
P
