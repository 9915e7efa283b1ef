for v in 0 1; do
echo "== RSLRL_SMEM_CARVEOUT=$v"
RSLRL_SMEM_CARVEOUT=$v timeout 200 python bench.py --steps 20 --warmup 5 --only-timed 2>&1 | tail -c 330
echo
RSLRL_SMEM_CARVEOUT=$v timeout 300 python profiles/experiments/c45_run.py c4 tf32 2>&1 | grep "C4"
done
