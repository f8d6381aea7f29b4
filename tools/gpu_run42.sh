set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for pdl in 1 0; do
VPM_TUNE_PDL=$pdl python bench.py --steps 100 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/bench_vp_pdl$pdl.json
for w in lb clb; do
VPM_TUNE_PDL=$pdl python bench.py --workload $w --steps 50 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/bench_${w}_pdl$pdl.json
done
python - <<P
import json
for w in ("vp","lb","clb"):
    d=json.load(open(f"gpurun_out/bench_{w}_pdl$pdl.json")); r=d["roofline"]
    print("RESULT pdl=$pdl", w, "ms/step %.4f"%d["ms_per_step"], "%.4e"%d["value"], "pass ms %.4f"%r["avg_launch_ms"])
P
done
VPM_TUNE_PDL=1 python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e --particles 1e4 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print('RESULT small pdl=1 us/step %.2f'%(1e3*d['ms_per_step']))"
VPM_TUNE_PDL=0 python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e --particles 1e4 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print('RESULT small pdl=0 us/step %.2f'%(1e3*d['ms_per_step']))"
