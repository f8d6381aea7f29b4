#!/bin/bash
# round 2, run Q (8 GPUs): the driver's default line at N = 8 (carried stagger, velocity-sorted LB / CLB, device-side rendezvous),
# N = 1 on the same box for the weak-scaling ratio, and the 100-step window
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 120 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-e2e --no-extras > gpurun_out/r2q_n1_samebox8_20.json 2>gpurun_out/r2q.err
timeout 120 python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu --no-e2e --no-extras > gpurun_out/r2q_n1_samebox8_100.json 2>>gpurun_out/r2q.err
timeout 500 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2q_n8_default.json 2>>gpurun_out/r2q.err
timeout 200 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --steps 100 --warmup 5 --no-cpu --no-e2e --no-extras > gpurun_out/r2q_n8_100.json 2>>gpurun_out/r2q.err
timeout 200 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu --no-e2e --no-extras > gpurun_out/r2q_n8_20.json 2>>gpurun_out/r2q.err
tail -4 gpurun_out/r2q.err
