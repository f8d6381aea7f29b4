#!/bin/bash
# round 2, run W (8 GPUs): fused all-reduce with all mailbox loads in flight -- N = 1 and N = 8 windows, LB / CLB lines at N = 8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 120 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-e2e --no-extras > gpurun_out/r2w_n1_samebox8_20.json 2>gpurun_out/r2w.err
timeout 200 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu --no-e2e --no-extras > gpurun_out/r2w_n8_20.json 2>>gpurun_out/r2w.err
timeout 200 $TR --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 --steps 100 --warmup 5 --no-cpu --no-e2e --no-extras > gpurun_out/r2w_n8_100.json 2>>gpurun_out/r2w.err
for w in lb clb; do
timeout 200 $TR --nproc-per-node 8 --master-port 29543 bench.py --gpus 8 --workload $w --steps 30 --warmup 5 --no-cpu --no-e2e --no-extras > gpurun_out/r2w_n8_$w.json 2>>gpurun_out/r2w.err
done
tail -3 gpurun_out/r2w.err
