for n in 1e3 1e4 1e5 2e5; do for w in lb clb; do for m in 0 2; do
VPM_TUNE_LBSORT=$m timeout 100 python bench.py --workload $w --particles $n --steps 200 --warmup 5 --no-cpu --no-e2e --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$n','$w','sort=$m', round(d['ms_per_step']*1e3,2),'us/step', d['gpu_launches'])"
done; done; done
