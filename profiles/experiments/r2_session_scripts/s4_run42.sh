timeout 120 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "loss" 2>&1 | tail -n 2
timeout 200 python profiles/experiments/kernel_table.py fp16 2>&1 | grep "kernel table\|ppo_loss\|mlp3\|dgrad_chain" | cut -c1-150
