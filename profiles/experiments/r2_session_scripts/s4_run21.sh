timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "linear" 2>&1 | tail -n 2
timeout 600 python tools/parity_report.py --out gpurun_out/parity_x3.json --precisions tf32x3 2>&1 | grep -v "^$" | cut -c1-700 | tail -n 4
timeout 200 python bench.py --steps 10 --warmup 3 --only-timed --precision tf32x3 2>&1 | tail -c 330
