timeout 120 python -m pytest tests/test_gpu_fp16_path.py -q -m gpu -x -k "dgrad_chain" 2>&1 | tail -n 12 | cut -c1-300
RSLRL_DGRAD_CHAIN=0 timeout 200 python bench.py --steps 20 --warmup 5 --only-timed 2>&1 | tail -c 330; echo
timeout 200 python bench.py --steps 20 --warmup 5 --only-timed 2>&1 | tail -c 330
