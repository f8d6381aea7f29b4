#!/bin/bash
# round 2, run K: LB stage 1 with the replicated table in the 256-worker ring (experimental build) vs the default build, twice each
mkdir -p gpurun_out
for rep in 1 2; do
python bench.py --workload lb --steps 30 --no-cpu --no-extras > gpurun_out/r2_k_lb_default_$rep.json 2>> gpurun_out/r2_k.err
VPM_B200_LIB=$PWD/vlasovparticlemethods.jl_b200/lib/libvpm_b200_exp.so python bench.py --workload lb --steps 30 --no-cpu --no-extras > gpurun_out/r2_k_lb_exp_$rep.json 2>> gpurun_out/r2_k.err
done
tail -3 gpurun_out/r2_k.err
python -c "
import json
for f in ('r2_k_lb_default_1','r2_k_lb_exp_1','r2_k_lb_default_2','r2_k_lb_exp_2'):
    w=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, round(w['ms_per_step'],4), {n:round(p['avg_launch_ms'],4) for n,p in w['passes'].items()})
"
