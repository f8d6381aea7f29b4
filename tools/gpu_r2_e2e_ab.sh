for i in 1 2 3; do timeout 120 python bench.py --steps 10 --warmup 5 --no-cpu --no-extras --e2e-steps 8 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('run$i', e['value'], e['pcie']['ms_per_step'], e['pcie']['copy_only_ms_per_step'], e['pcie']['frac_of_pcie_ceiling'])"
done
VPM_TUNE_E2E_CHUNKS=1 timeout 120 python bench.py --steps 10 --warmup 5 --no-cpu --no-extras --e2e-steps 8 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('chunks1', e['value'], e['pcie']['ms_per_step'], e['pcie']['copy_only_ms_per_step'], e['pcie']['frac_of_pcie_ceiling'])"
