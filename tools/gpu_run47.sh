set -x
timeout 900 python -m pytest tests/test_gpu_uniform_weight.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 100 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print('RESULT vp ms/step %.4f'%d['ms_per_step'], 'uw', d['uniform_weight_variant'])"
