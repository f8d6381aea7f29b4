set -x
for nb in 16 32 48 64 100 128; do
  python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e --n-basis $nb 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); r=d['roofline']; print('NB', $nb, 'ms/step %.4f'%d['ms_per_step'], 'pass_ms %.4f'%r['avg_launch_ms'], 'GB/s %.0f'%r['achieved'], 'field %.3f'%r['field_kernel_share'])"
done
python bench.py --workload lb --steps 20 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); r=d['roofline']; print('LB ms/step %.4f'%d['ms_per_step'], 'GB/s %.0f'%r['achieved'])"
python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_vp_d.json; cat gpurun_out/bench_vp_d.json
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
