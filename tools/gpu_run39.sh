set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 100 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/bench_vp_h.json
for w in lb clb; do
python bench.py --workload $w --steps 50 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/bench_${w}_h.json
done
python - <<'P'
import json
for w in ("vp","lb","clb"):
    d=json.load(open(f"gpurun_out/bench_{w}_h.json")); r=d["roofline"]
    print("RESULT", w, "ms/step %.4f"%d["ms_per_step"], "%.4e"%d["value"], "pass ms %.4f"%r["avg_launch_ms"], "field share %.4f"%r["field_kernel_share"])
P
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_clb_h.csv python bench.py --workload clb --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch_clb.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches_vp_h.csv python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch_vp.log 2>&1
python - <<'P'
import csv
for f in ("clb","vp"):
    rows=[r for r in csv.reader(open(f"gpurun_out/launches_{f}_h.csv")) if len(r)>5 and r[0].isdigit()]
    for r in rows[-10:]:
        if 'field' in r[4]: print(f, r[4][:50], r[-1])
P
