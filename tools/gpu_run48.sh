set -x
for t in 1 4; do
VPM_TUNE_TMA=$t python bench.py --steps 100 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); r=d['roofline']; print('RESULT vp tma=$t ms/step %.4f pass %.4f GB/s %.0f'%(d['ms_per_step'], r['avg_launch_ms'], r['achieved']))"
done
