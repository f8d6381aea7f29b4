set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for nb in 16 64 100 128 256 512 1024 2048; do
  python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e --n-basis $nb 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); r=d['roofline']; print('NB', $nb, 'ms/step %.4f'%d['ms_per_step'], 'pass_ms %.4f'%r['avg_launch_ms'], 'GB/s %.0f'%r['achieved'], 'field %.3f'%r['field_kernel_share'])"
done
VPM_TUNE_HM=3 python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e --n-basis 16 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); r=d['roofline']; print('TILED NB 16', 'ms/step %.4f'%d['ms_per_step'], 'pass_ms %.4f'%r['avg_launch_ms'])"
VPM_TUNE_HM=3 python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e --n-basis 64 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); r=d['roofline']; print('TILED NB 64', 'ms/step %.4f'%d['ms_per_step'], 'pass_ms %.4f'%r['avg_launch_ms'])"
