timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "rnd_reward" 2>&1 | grep -v "^$" | tail -n 40 | cut -c1-300
timeout 200 python profiles/experiments/host_step_breakdown.py device 2>&1 | tail -n 2
timeout 200 python profiles/experiments/host_step_breakdown.py host 2>&1 | tail -n 2
