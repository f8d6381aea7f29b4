set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_vp_e.json; cat gpurun_out/bench_vp_e.json
