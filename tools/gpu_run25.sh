set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for t in 1 0; do
VPM_TUNE_TMA=$t python bench.py --steps 100 --warmup 5 --no-cpu --no-e2e --field frozen 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); r=d['roofline']; print('FROZEN TMA', $t, 'ms/step %.4f'%d['ms_per_step'], 'pass_ms %.4f'%r['avg_launch_ms'], 'GB/s %.0f'%r['achieved'])"
done
