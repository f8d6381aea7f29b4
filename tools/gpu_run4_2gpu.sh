set -x
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 2>&1 | tail -3 > gpurun_out/bench_vp_n2.json; cat gpurun_out/bench_vp_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --workload clb 2>&1 | tail -3 > gpurun_out/bench_clb_n2.json; cat gpurun_out/bench_clb_n2.json
