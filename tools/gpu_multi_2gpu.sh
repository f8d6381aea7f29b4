set -x
timeout 300 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -2
for w in vp lb; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --workload $w --steps 50 --warmup 5 --no-e2e --no-cpu 2>&1 | tail -1 > gpurun_out/bench_${w}_n2_pdl.json
python -c "
import json; d=json.load(open('gpurun_out/bench_${w}_n2_pdl.json')); print('RESULT $w n=2 ms/step %.4f %.4e'%(d['ms_per_step'], d['value']), d['config']['parallelism'])"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 50 --warmup 5 --no-e2e --no-cpu --comm nccl 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print('RESULT vp nccl n=2 ms/step %.4f'%d['ms_per_step'])"
