set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for np in 1e4 1e5 1e6 1e7; do
  python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e --particles $np 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); r=d['roofline']; print('N', '$np', 'ms/step %.5f'%d['ms_per_step'], 'value %.3e'%d['value'])"
done
