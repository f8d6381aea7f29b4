set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 100 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/bench_vp_g.json; python -c "
import json; d=json.load(open('gpurun_out/bench_vp_g.json')); print('RESULT vp ms/step %.4f  %.3e p-steps/s  pass ms %.4f GB/s %.0f uw %s'%(d['ms_per_step'], d['value'], d['roofline']['avg_launch_ms'], d['roofline']['achieved'], d['uniform_weight_variant']))"
