set -x
timeout 300 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -2
for w in lb clb; do
timeout 300 python bench.py --gpus 1 --workload $w --steps 50 --warmup 5 --no-e2e --no-cpu 2>&1 | tail -1 > gpurun_out/r1_bench_${w}_n1_samebox8.json
for n in 8 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2964$n bench.py --gpus $n --workload $w --steps 50 --warmup 5 --no-e2e --no-cpu 2>&1 | tail -1 > gpurun_out/r1_bench_${w}_n${n}_p2p.json
done
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --workload clb --steps 50 --warmup 5 --no-e2e --no-cpu --comm nccl 2>&1 | tail -1 > gpurun_out/r1_bench_clb_n8_nccl.json
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/r1_bench_*lb_n*")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], "ms/step %.4f"%d["ms_per_step"], "%.4e"%d["value"], d["config"]["parallelism"])
    except Exception as e: print(f, "ERR", e, open(f).read()[-300:])
P
