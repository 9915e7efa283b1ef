RSLRL_HOST_PROFILE=1 timeout 300 python profiles/experiments/c45_run.py c4 tf32 2>&1 | grep -v Warn | cut -c1-200 | head -50
