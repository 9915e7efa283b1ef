set -x
RSLRL_KERNEL_TABLE=1 RSLRL_HOST_PROFILE=1 RSLRL_C45_WARMUP=3 RSLRL_C45_ITERS=3 python profiles/experiments/c45_run.py c4 tf32 > gpurun_out/s4_c4_table.log 2>&1
RSLRL_KERNEL_TABLE=1 RSLRL_C45_WARMUP=4 RSLRL_C45_ITERS=3 python profiles/experiments/c45_run.py c5 tf32 > gpurun_out/s4_c5_table.log 2>&1
python -m rsl_rl_b200._build --out rsl_rl_b200/lib/librslrl_b200_trace.so -DRSLRL_GAE_TRACE -DRSLRL_AR_TRACE > gpurun_out/s4_tracebuild.log 2>&1
RSLRL_B200_LIB=rsl_rl_b200/lib/librslrl_b200_trace.so python profiles/experiments/gae_trace.py > gpurun_out/s4_gae_trace.log 2>&1
tail -n 5 gpurun_out/s4_gae_trace.log
