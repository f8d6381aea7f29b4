set -x
ncu --set full --clock-control none --import-source on -k regex:lb_pass -s 9 -c 8 -o gpurun_out/prof_clb_tma python bench.py --workload clb --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_clb_tma.log 2>&1
VPM_TUNE_LBTMA=0 ncu --set full --clock-control none --import-source on -k regex:lb_pass -s 9 -c 8 -o gpurun_out/prof_clb_reg python bench.py --workload clb --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_clb_reg.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lb_pass -s 5 -c 4 -o gpurun_out/prof_lb_tma python bench.py --workload lb --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_lb_tma.log 2>&1
ls -la gpurun_out/*.ncu-rep
