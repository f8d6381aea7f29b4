set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in lb clb; do
python bench.py --workload $w --steps 30 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/bench_${w}_k.json; cat gpurun_out/bench_${w}_k.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_${w}_k.csv python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch_${w}.log 2>&1
done
python - <<'P'
import csv
for w in ("lb","clb"):
    rows=[r for r in csv.reader(open(f"gpurun_out/launches_{w}_k.csv")) if len(r)>5 and r[0].isdigit()]
    for r in rows[-14:]: print(w, r[4][:60], r[-1])
P
