#!/bin/bash
# round 2, 2 GPUs: multi-rank parity tests (fused p2p all-reduce and NCCL) with the log kept, bench at N=2 with parity_check
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_2gpu.txt 2>&1
python -m pytest tests/test_gpu_multi.py -v 2>&1 | tail -25 > gpurun_out/r2_multi_2gpu_pytest.log
cat gpurun_out/r2_multi_2gpu_pytest.log | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_p2p.json 2> gpurun_out/r2_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --comm nccl --no-e2e > gpurun_out/r2_bench_n2_nccl.json 2>> gpurun_out/r2_bench_n2.err
tail -5 gpurun_out/r2_bench_n2.err
python -c "
import json
for f in ('r2_bench_n2_p2p','r2_bench_n2_nccl'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['comm'], d['parity_check'])
"
