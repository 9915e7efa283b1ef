set -x
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -n 4 | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -i "smoke\|error\|assert" | cut -c1-300
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench_fp16.json 2> gpurun_out/r2_final_bench_fp16.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err
python bench.py --steps 10 --warmup 3 --precision tf32x3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r2_final_bench_tf32x3.json 2> gpurun_out/r2_final_bench_tf32x3.err
python bench.py --steps 5 --warmup 3 --config c3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r2_final_bench_c3.json 2> gpurun_out/r2_final_bench_c3.err
python profiles/experiments/c45_run.py c4 tf32 > gpurun_out/r2_final_c4.log 2>&1
python profiles/experiments/c45_run.py c5 fp16 > gpurun_out/r2_final_c5.log 2>&1
bash profiles/experiments/prof_r2.sh > gpurun_out/r2_prof.log 2>&1
tail -n 2 gpurun_out/r2_final_c4.log gpurun_out/r2_final_c5.log
