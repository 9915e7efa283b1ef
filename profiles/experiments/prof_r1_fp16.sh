# round-1 profile set of the fp16-storage build (run under gpurun from the repo root)
set -x
CMD="python bench.py --steps 1 --warmup 3 --only-timed"
$CMD > gpurun_out/p4_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/p4_launches_fp16.csv $CMD > gpurun_out/p4_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"mlp3_forward|tc_gemm|ppo_loss|allreduce_clip_adam|head_backward" -s 60 -c 9 -o gpurun_out/p4_update_fp16 $CMD > gpurun_out/p4_ncu2.log 2>&1
tail -2 gpurun_out/p4_plain.log gpurun_out/p4_ncu2.log
