set -x
for nb in 16 32 48 64 128 512 2048; do
  python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e --n-basis $nb 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); r=d['roofline']; print('NB', $nb, 'ms/step %.4f'%d['ms_per_step'], 'pass_ms %.4f'%r['avg_launch_ms'], 'GB/s %.0f'%r['achieved'], 'field %.3f'%r['field_kernel_share'])"
done
for k in 2 3 5 6; do
  python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e --order $k 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); r=d['roofline']; print('ORDER', $k, 'ms/step %.4f'%d['ms_per_step'], 'pass_ms %.4f'%r['avg_launch_ms'], 'GB/s %.0f'%r['achieved'])"
done
for np in 1e4 1e5 1e6 1e7; do
  python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e --particles $np 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); r=d['roofline']; print('N', '$np', 'ms/step %.5f'%d['ms_per_step'], 'value %.3e'%d['value'])"
done
