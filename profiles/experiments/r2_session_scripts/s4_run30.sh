timeout 120 python -m pytest tests/test_gpu_fp16_path.py -q -m gpu -x -k "dgrad_chain" 2>&1 | tail -n 12 | cut -c1-300
timeout 200 python bench.py --steps 20 --warmup 5 --only-timed 2>&1 | tail -c 330; echo
timeout 200 python profiles/experiments/kernel_table.py fp16 2>&1 | grep "dgrad_chain\|tc_gemm\|mlp3" | cut -c1-150
