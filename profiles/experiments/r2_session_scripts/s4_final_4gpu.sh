timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline --no-reference-cuda > gpurun_out/r2_final_bench_4gpu.json 2> gpurun_out/r2_final_bench_4gpu.err
tail -c 200 gpurun_out/r2_final_bench_4gpu.json
