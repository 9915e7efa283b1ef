set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
CMD="python profiles/experiments/c3_run.py tf32 65536 1"
$CMD > gpurun_out/c3_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/c3_launches.csv $CMD > gpurun_out/c3_ncu.log 2>&1
tail -2 gpurun_out/c3_plain.log
python bench.py --no-cpu-baseline > gpurun_out/c3_bench_c2.json 2>/dev/null
