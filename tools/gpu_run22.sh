set -x
VPM_TUNE_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "strang or large_prop" 2>&1 | tail -3
for t in 1 2 3; do
  VPM_TUNE_TMA=$t python bench.py --steps 100 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); r=d['roofline']; print('TMA', $t, 'ms/step %.4f'%d['ms_per_step'], 'pass_ms %.4f'%r['avg_launch_ms'], 'GB/s %.0f'%r['achieved'], 'uw', d['uniform_weight_variant']['ms_per_step'])"
done
