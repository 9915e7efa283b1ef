set -x
python bench.py --steps 20 --warmup 5 --only-timed > gpurun_out/s4_head1.json 2> gpurun_out/s4_head1.err
RSLRL_FUSED_HEAD_DGRAD=0 python bench.py --steps 20 --warmup 5 --only-timed > gpurun_out/s4_head0.json 2> gpurun_out/s4_head0.err
python profiles/experiments/c45_run.py c4 tf32 > gpurun_out/s4_c4.log 2>&1
python profiles/experiments/c45_run.py c5 tf32 > gpurun_out/s4_c5.log 2>&1
python profiles/experiments/c45_run.py c5 fp16 > gpurun_out/s4_c5_fp16.log 2>&1
tail -n 3 gpurun_out/s4_*.log gpurun_out/s4_*.json
