set -x
nvidia-smi -L
nvidia-smi topo -m | head -12
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -15
for comm in p2p nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 100 --warmup 5 --comm $comm --no-e2e 2>&1 | tail -2 > gpurun_out/bench_vp_n2_$comm.json; cat gpurun_out/bench_vp_n2_$comm.json
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 3 --workload clb --comm p2p 2>&1 | tail -2 > gpurun_out/bench_clb_n2_p2p.json; cat gpurun_out/bench_clb_n2_p2p.json
timeout 600 python -m pytest tests/test_gpu_physics.py -x -q 2>&1 | tail -5
