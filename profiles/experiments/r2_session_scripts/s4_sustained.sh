timeout 280 python bench.py --steps 2000 --warmup 20 --no-cpu-baseline --no-reference-cuda > gpurun_out/r2_bench_fp16_2000steps.json 2> gpurun_out/r2_bench_fp16_2000steps.err
tail -c 300 gpurun_out/r2_bench_fp16_2000steps.json
