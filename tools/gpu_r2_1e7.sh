# per-step cost at 1e7 particles per GPU (the lower end of BASELINE config 3): how much of the step are the single-CTA field kernels?
for w in vp lb clb; do
timeout 100 python bench.py --workload $w --particles 1e7 --steps 200 --warmup 5 --no-cpu --no-e2e --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$w', 'ms/step', round(d['ms_per_step'],4), 'value %.4g'%d['value'], 'field share', round(r['field_kernel_share'],3), 'field avg us', round(r.get('field_kernel_avg_ms',0)*1e3,1), 'pass avg', r.get('avg_launch_ms'), {k:round(v['avg_launch_ms'],4) for k,v in d.get('passes',{}).items()} if isinstance(d.get('passes'),dict) else '')"
done
