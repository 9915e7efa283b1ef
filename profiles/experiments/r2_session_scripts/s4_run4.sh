set -x
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/s4_t4.log 2>&1; tail -n 5 gpurun_out/s4_t4.log
python -m pytest tests/test_gpu_ppo_parity.py -x -q -m gpu -k "golden or c1_shape" > gpurun_out/s4_t4b.log 2>&1; tail -n 3 gpurun_out/s4_t4b.log
python -m rsl_rl_b200._build --out rsl_rl_b200/lib/librslrl_b200_trace.so -DRSLRL_GAE_TRACE -DRSLRL_AR_TRACE > gpurun_out/s4_tracebuild.log 2>&1
RSLRL_B200_LIB=rsl_rl_b200/lib/librslrl_b200_trace.so python profiles/experiments/gae_trace.py > gpurun_out/s4_gae_trace2.log 2>&1
tail -n 3 gpurun_out/s4_gae_trace2.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-cuda > gpurun_out/s4_bench4.json 2> gpurun_out/s4_bench4.err
