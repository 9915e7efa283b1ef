set -x
timeout 300 python -m pytest tests/test_gpu_ppo_parity.py -x -q -m gpu -s -k "3xtf32" 2>&1 | grep "tf32x3 C1\|passed\|failed" 
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "linear" 2>&1 | tail -n 2
timeout 200 python bench.py --steps 10 --warmup 3 --only-timed --precision tf32x3 2>&1 | tail -c 400
timeout 300 python -m pytest tests/test_gpu_composite_parity.py -x -q -m gpu 2>&1 | tail -n 5
RSLRL_KERNEL_TABLE=1 timeout 300 python profiles/experiments/c45_run.py c5 fp16 2>&1 | grep -v Warn | cut -c1-180 | head -24
timeout 300 python profiles/experiments/gae_modes.py 2>&1 | grep "mode\|floor"
