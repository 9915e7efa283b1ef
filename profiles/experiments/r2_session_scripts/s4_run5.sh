python profiles/experiments/gae_modes.py 2>&1 | grep -v Warn
RSLRL_GAE_TMA=0 python profiles/experiments/gae_modes.py 2>&1 | grep "mode"
RSLRL_GAE_CLUSTER_MAX=0 python profiles/experiments/gae_modes.py 2>&1 | grep "mode"
RSLRL_GAE_CLUSTER_MAX=8 python profiles/experiments/gae_modes.py 2>&1 | grep "mode"
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "loss" 2>&1 | tail -n 3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-cuda > gpurun_out/s4_bench5.json 2> gpurun_out/s4_bench5.err
