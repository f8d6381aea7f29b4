set -x
timeout 600 python tools/physics_checks.py 1e8 > gpurun_out/r1_physics_1e8.json 2> gpurun_out/physics.err; tail -2 gpurun_out/physics.err; cut -c1-900 gpurun_out/r1_physics_1e8.json
