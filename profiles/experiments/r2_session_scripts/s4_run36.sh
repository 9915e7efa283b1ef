python -m rsl_rl_b200._build --out rsl_rl_b200/lib/librslrl_b200_trace.so -DRSLRL_DC_TRACE > gpurun_out/s4_tracebuild2.log 2>&1
RSLRL_B200_LIB=rsl_rl_b200/lib/librslrl_b200_trace.so timeout 120 python profiles/experiments/dgrad_chain_trace.py 2>&1 | tail -n 14
