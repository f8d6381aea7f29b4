set -x
nvidia-smi -L | head -8
nproc
for n in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 100 --warmup 5 2>&1 | tail -2 > gpurun_out/bench_vp_n${n}_p2p.json; cat gpurun_out/bench_vp_n${n}_p2p.json
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 100 --warmup 5 --comm nccl --no-e2e 2>&1 | tail -2 > gpurun_out/bench_vp_n8_nccl.json; cat gpurun_out/bench_vp_n8_nccl.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 20 --warmup 3 --workload clb 2>&1 | tail -2 > gpurun_out/bench_clb_n8_p2p.json; cat gpurun_out/bench_clb_n8_p2p.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus 8 --steps 10 --warmup 1 2>&1 | tail -2
