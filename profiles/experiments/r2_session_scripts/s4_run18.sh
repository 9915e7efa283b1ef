timeout 300 python -m pytest tests/test_gpu_ppo_parity.py -q -m gpu -x -k "host_resident or golden" 2>&1 | tail -n 5 | cut -c1-300
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -i "smoke\|error\|assert" | cut -c1-300
RSLRL_HOST_ACTIONS=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-cuda 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('HOST_ACTIONS=0', d['ms_per_step'], d['e2e']['ms_per_step'], d['phases_ms'])"
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-cuda 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('HOST_ACTIONS=1', d['ms_per_step'], d['e2e']['ms_per_step'], d['phases_ms'])"
timeout 200 python profiles/experiments/host_step_breakdown.py host 2>&1 | tail -n 1
