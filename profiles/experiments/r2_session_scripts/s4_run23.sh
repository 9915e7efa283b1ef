timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-reference-cuda > gpurun_out/r2_final_bench_8gpu.json 2> gpurun_out/r2_final_bench_8gpu.err
tail -c 400 gpurun_out/r2_final_bench_8gpu.json
