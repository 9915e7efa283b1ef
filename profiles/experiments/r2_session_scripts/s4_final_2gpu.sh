timeout 600 python -m pytest tests/test_gpu_multi_rank.py -q -m gpu 2>&1 | tail -n 3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-reference-cuda > gpurun_out/r2_final_bench_2gpu.json 2> gpurun_out/r2_final_bench_2gpu.err
tail -c 200 gpurun_out/r2_final_bench_2gpu.json
