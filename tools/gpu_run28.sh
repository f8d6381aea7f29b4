set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in lb clb; do
python bench.py --workload $w --steps 30 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/bench_${w}_t.json; cat gpurun_out/bench_${w}_t.json
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_clb_t.csv python bench.py --workload clb --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch_clb.log 2>&1
python - <<'P'
import csv
for w in ("clb",):
    rows=[r for r in csv.reader(open(f"gpurun_out/launches_{w}_t.csv")) if len(r)>5 and r[0].isdigit()]
    for r in rows[-18:]: print(w, r[4][:70], r[-1])
P
