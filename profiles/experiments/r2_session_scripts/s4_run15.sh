timeout 600 python -m pytest tests/test_gpu_composite_parity.py -q -m gpu -x -k "recurrent or lstm" 2>&1 | tail -n 4 | cut -c1-300
RSLRL_HOST_PROFILE=1 timeout 300 python profiles/experiments/c45_run.py c4 tf32 2>&1 | grep -v Warn | cut -c1-200 | head -16
RSLRL_RECURRENT_GRAPH=0 timeout 300 python profiles/experiments/c45_run.py c4 tf32 2>&1 | grep "C4"
