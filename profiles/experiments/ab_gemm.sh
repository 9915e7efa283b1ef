# A/B of GEMM kernel variants: per-layer microbench for every rsl_rl_b200/lib/var_*.so, in-kernel timeline for the traced ones
for so in rsl_rl_b200/lib/var_*.so; do
  echo "=== $so"
  RSLRL_B200_LIB=$PWD/$so python profiles/experiments/gemm_microbench.py tf32 2>&1 | tail -7
done
for so in $TRACE_LIBS; do
  for k in f d w; do
    echo "=== trace $so $k"
    RSLRL_B200_LIB=$PWD/$so python profiles/experiments/gemm_trace.py $k 2>&1 | tail -20
  done
done
