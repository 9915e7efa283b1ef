# round-2 profile set of the fp16-storage build (run under gpurun from the repo root; outputs in gpurun_out/)
set -x
CMD="python bench.py --steps 1 --warmup 3 --only-timed"
$CMD > gpurun_out/r2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_fp16.csv $CMD > gpurun_out/r2_ncu1.log 2>&1
# one minibatch of update(): the 3 dense launches (skip the 25 rollout / last-value forwards of the timed iteration)
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"mlp3_forward|tc_gemm|dgrad_chain" -s 25 -c 3 -o gpurun_out/r2_dense_fp16 $CMD > gpurun_out/r2_ncu2.log 2>&1
# the HBM-side kernels of the same iteration: one launch each
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"ppo_loss|allreduce_clip_adam|gae_norm|permute_gather" -c 6 -o gpurun_out/r2_hbm_fp16 $CMD > gpurun_out/r2_ncu3.log 2>&1
for f in gpurun_out/r2_plain.log gpurun_out/r2_ncu2.log gpurun_out/r2_ncu3.log; do tail -n 2 $f; done
