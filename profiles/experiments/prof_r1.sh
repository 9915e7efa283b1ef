set -x
CMD="python bench.py --steps 1 --warmup 3 --only-timed"
$CMD > gpurun_out/p1_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/p1_launches.csv $CMD > gpurun_out/p1_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"tc_gemm|ppo_loss|adam_step|grad_norm|head" -s 100 -c 16 -o gpurun_out/p1_update $CMD > gpurun_out/p1_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gae_norm|permute_gather|store_transition|gaussian_act" -c 4 -o gpurun_out/p1_small $CMD > gpurun_out/p1_ncu3.log 2>&1
tail -3 gpurun_out/p1_plain.log gpurun_out/p1_ncu2.log gpurun_out/p1_ncu3.log
