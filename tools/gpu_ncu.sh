#!/bin/bash
# One ncu --set full capture of every kernel of one timed users range (512 users = one chunk: the per-chunk launch shapes of
# the full step) plus the joint / training kernels' standalone launches; raw-page CSVs come back, the .ncu-rep stays.
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/prof_users \
    python bench.py --steps 1 --warmup 1 --pool-items 131072 --users-per-gpu 512 --no-cpu-baseline --train-batch 0 --profile-range users \
    > gpurun_out/prof_users.out 2>&1
ncu -i /tmp/prof_users.ncu-rep --page raw --csv > gpurun_out/prof_users_raw.csv 2> gpurun_out/prof_users_raw.err
for k in gemm_bf16_cg2 attention_tc; do
  ncu -i /tmp/prof_users.ncu-rep --page source --csv -k regex:$k -c 1 > gpurun_out/src_users_$k.csv 2>> gpurun_out/prof_users_raw.err
done
timeout 600 ncu --set full --clock-control none -k regex:'list_scores|infonce|inject_tokens|layernorm_bwd|reconstruction' -c 12 -f -o /tmp/prof_side \
    python tools/gpu_ncu_side.py > gpurun_out/prof_side.out 2>&1
ncu -i /tmp/prof_side.ncu-rep --page raw --csv > gpurun_out/prof_side_raw.csv 2>> gpurun_out/prof_users_raw.err
ls -la /tmp/*.ncu-rep gpurun_out
