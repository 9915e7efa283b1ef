timeout 600 python -m pytest tests/test_gpu_ppo_parity.py tests/test_gpu_kernels.py tests/test_gpu_composite_parity.py -q -m gpu -x 2>&1 | tail -n 4 | cut -c1-300
timeout 200 python profiles/experiments/host_step_breakdown.py device 2>&1 | tail -n 1
timeout 200 python profiles/experiments/host_step_breakdown.py host 2>&1 | tail -n 1
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-cuda 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['ms_per_step'], d['e2e']['ms_per_step'], d['phases_ms'])"
