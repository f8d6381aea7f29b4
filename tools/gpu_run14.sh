set -x
for t in 0 1 2; do for nb in 128 512; do
  VPM_TUNE_TILE=$t python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e --n-basis $nb 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); r=d['roofline']; print('TILE', $t, 'NB', $nb, 'ms/step %.4f'%d['ms_per_step'], 'pass_ms %.4f'%r['avg_launch_ms'], 'GB/s %.0f'%r['achieved'])"
done; done
