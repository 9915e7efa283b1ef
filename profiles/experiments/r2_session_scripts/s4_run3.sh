set -x
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fp16_path.py -x -q -m gpu > gpurun_out/s4_t3.log 2>&1; tail -n 5 gpurun_out/s4_t3.log
python -m pytest tests/test_gpu_ppo_parity.py -x -q -m gpu -k "fp16 or tensor_core" > gpurun_out/s4_t3b.log 2>&1; tail -n 5 gpurun_out/s4_t3b.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-cuda > gpurun_out/s4_bench3.json 2> gpurun_out/s4_bench3.err
RSLRL_FUSED_HEAD_DGRAD=0 python bench.py --steps 20 --warmup 5 --only-timed > gpurun_out/s4_bench3_head0.json 2>&1
tail -c 600 gpurun_out/s4_bench3_head0.json
