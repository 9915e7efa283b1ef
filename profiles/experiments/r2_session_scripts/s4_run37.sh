python -m rsl_rl_b200._build --out rsl_rl_b200/lib/librslrl_b200_w4g5.so -DRSLRL_MF_W_STAGES=4 -DRSLRL_MF_GRANULES=5 > gpurun_out/s4_variantbuild.log 2>&1
RSLRL_B200_LIB=rsl_rl_b200/lib/librslrl_b200_w4g5.so timeout 120 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "mlp3" 2>&1 | tail -n 2
RSLRL_B200_LIB=rsl_rl_b200/lib/librslrl_b200_w4g5.so timeout 200 python bench.py --steps 20 --warmup 5 --only-timed 2>&1 | tail -c 330; echo
RSLRL_B200_LIB=rsl_rl_b200/lib/librslrl_b200_w4g5.so timeout 200 python profiles/experiments/kernel_table.py fp16 2>&1 | grep "mlp3" | cut -c1-150
timeout 200 python bench.py --steps 20 --warmup 5 --only-timed 2>&1 | tail -c 330; echo
