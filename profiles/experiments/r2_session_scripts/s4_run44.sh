for i in 1 2; do
RSLRL_FUSED_HEAD_DGRAD=0 timeout 200 python bench.py --steps 40 --warmup 5 --only-timed 2>&1 | tail -c 330 | grep -o '"ms_per_step": [0-9.]*' | sed 's/^/HEAD=0 /'
RSLRL_FUSED_HEAD_DGRAD=1 timeout 200 python bench.py --steps 40 --warmup 5 --only-timed 2>&1 | tail -c 330 | grep -o '"ms_per_step": [0-9.]*' | sed 's/^/HEAD=1 /'
done
