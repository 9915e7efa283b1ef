set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "linear" > gpurun_out/s4_t6.log 2>&1; tail -n 15 gpurun_out/s4_t6.log
timeout 300 python -m pytest tests/test_gpu_ppo_parity.py -x -q -m gpu -s -k "3xtf32" > gpurun_out/s4_t6b.log 2>&1; tail -n 15 gpurun_out/s4_t6b.log
timeout 200 python bench.py --steps 10 --warmup 3 --only-timed --precision tf32x3 > gpurun_out/s4_bench6_x3.json 2>&1; tail -c 500 gpurun_out/s4_bench6_x3.json
timeout 200 python bench.py --steps 10 --warmup 3 --only-timed --precision fp32 > gpurun_out/s4_bench6_fp32.json 2>&1; tail -c 500 gpurun_out/s4_bench6_fp32.json
RSLRL_LOSS_QUAD_MAX_ROWS=100000000 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/s4_bench6_quad.json 2> gpurun_out/s4_bench6_quad.err
