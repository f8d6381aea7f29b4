#!/bin/bash
# round 2, 8 GPUs: default bench line at N=8 (fused p2p all-reduce) with parity_check, and the NCCL transport for comparison
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_8gpu.txt 2>&1
(lscpu | head -25; numactl --hardware 2>&1 | head -12) > gpurun_out/r2_host_8gpu.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8_p2p.json 2> gpurun_out/r2_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 100 --warmup 5 --no-e2e --no-extras > gpurun_out/r2_bench_n8_p2p_100.json 2>> gpurun_out/r2_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 100 --warmup 5 --comm nccl --no-e2e --no-extras > gpurun_out/r2_bench_n8_nccl_100.json 2>> gpurun_out/r2_bench_n8.err
python bench.py --steps 100 --warmup 5 --no-e2e --no-extras --no-cpu > gpurun_out/r2_bench_n1_samebox8_100.json 2>> gpurun_out/r2_bench_n8.err
tail -5 gpurun_out/r2_bench_n8.err
python -c "
import json
for f in ('r2_bench_n8_p2p','r2_bench_n8_p2p_100','r2_bench_n8_nccl_100','r2_bench_n1_samebox8_100'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['comm'], d['roofline']['field_kernel_share'], (d.get('parity_check') or {}).get('ok'))
"
