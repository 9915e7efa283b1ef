timeout 800 python -m pytest tests -q -m gpu 2>&1 | tail -n 12
RSLRL_KERNEL_TABLE=1 timeout 300 python profiles/experiments/c45_run.py c5 fp16 2>&1 | grep -v Warn | cut -c1-180 | head -14
timeout 300 python profiles/experiments/c45_run.py c4 tf32 2>&1 | grep "C4"
