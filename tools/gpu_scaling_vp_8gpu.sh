set -x
for n in 8 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2963$n bench.py --gpus $n --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_vp_n${n}_p2p.json; cat gpurun_out/bench_vp_n${n}_p2p.json
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 --steps 100 --warmup 5 --comm nccl --no-e2e 2>&1 | tail -1 > gpurun_out/bench_vp_n8_nccl.json; cat gpurun_out/bench_vp_n8_nccl.json
timeout 400 python bench.py --gpus 1 --steps 100 --warmup 5 --no-e2e --no-cpu 2>&1 | tail -1 > gpurun_out/bench_vp_n1_samebox.json; cat gpurun_out/bench_vp_n1_samebox.json
timeout 300 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -2
