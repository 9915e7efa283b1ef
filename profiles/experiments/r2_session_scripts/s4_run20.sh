timeout 400 python -m pytest tests/test_gpu_ppo_parity.py -q -m gpu -x -s -k "3xtf32" 2>&1 | grep -v "^$" | tail -n 12 | cut -c1-600
timeout 600 python tools/parity_report.py --out gpurun_out/parity_r2.json 2>&1 | tail -n 3 | cut -c1-300
