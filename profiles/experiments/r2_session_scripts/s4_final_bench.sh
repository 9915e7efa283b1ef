python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench_fp16.json 2> gpurun_out/r2_final_bench_fp16.err
timeout 200 python -m pytest tests/test_gpu_ppo_parity.py tests/test_gpu_fp16_path.py -q -m gpu -x -k "fp16" 2>&1 | tail -n 2
