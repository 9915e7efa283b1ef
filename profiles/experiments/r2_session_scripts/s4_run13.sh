timeout 600 python -m pytest tests/test_gpu_composite_parity.py -q -m gpu -x 2>&1 | tail -n 30 | cut -c1-300
RSLRL_KERNEL_TABLE=1 timeout 300 python profiles/experiments/c45_run.py c4 tf32 2>&1 | grep -v Warn | cut -c1-180 | head -12
RSLRL_RECURRENT_GRAPH=0 timeout 300 python profiles/experiments/c45_run.py c4 tf32 2>&1 | grep "C4"
timeout 300 python profiles/experiments/c45_run.py c4 tf32x3 2>&1 | grep "C4"
