RSLRL_GEMM_PRECISION=tf32x3 timeout 800 python -m pytest tests -q -m gpu -x 2>&1 | tail -n 25
