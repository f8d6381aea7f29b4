set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/lb_checks.py 1e7 50000 > gpurun_out/r1_physics_clb_1e7.json 2> gpurun_out/lb_checks.err; tail -2 gpurun_out/lb_checks.err; cut -c1-600 gpurun_out/r1_physics_clb_1e7.json
