#!/bin/bash
# One gpurun call: GPU tests, full bench, ncu launch lists and full captures of the top kernels.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh [tests] [bench] [launches] [ncu]
set -u
mkdir -p gpurun_out
what="${*:-tests bench launches ncu}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for w in $what; do
case $w in
tests)
  timeout 600 python -m pytest tests -m gpu -q --timeout 200 --durations=12 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  timeout 150 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
  ;;
bench)
  timeout 700 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?" >> gpurun_out/bench_full.err
  ;;
launches)
  # small pool so the run is short; every launch of one timed user step / two item chunks (cold-cache, serialised)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches_users.csv \
      python bench.py --steps 1 --warmup 1 --pool-items 131072 --users-per-gpu 1024 --no-cpu-baseline --train-batch 0 --profile-range users \
      > gpurun_out/launches_users.out 2>&1
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches_items.csv \
      python bench.py --steps 1 --warmup 1 --pool-items 12288 --users-per-gpu 128 --no-cpu-baseline --train-batch 0 --profile-range items \
      > gpurun_out/launches_items.out 2>&1
  ;;
ncu)
  # one --set full capture of EVERY kernel of one user chunk (256 users: the per-chunk shapes are those of the full
  # step) and of one item batch; the .ncu-rep stays on the box (too large), raw-page CSVs come back
  timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -f -o /tmp/prof_users \
      python bench.py --steps 1 --warmup 1 --pool-items 131072 --users-per-gpu 512 --no-cpu-baseline --train-batch 0 --profile-range users \
      > gpurun_out/prof_users.out 2>&1
  ncu -i /tmp/prof_users.ncu-rep --page raw --csv > gpurun_out/prof_users_raw.csv 2> gpurun_out/prof_users_raw.err
  timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -f -o /tmp/prof_items \
      python bench.py --steps 1 --warmup 1 --pool-items 8192 --users-per-gpu 128 --no-cpu-baseline --train-batch 0 --profile-range items \
      > gpurun_out/prof_items.out 2>&1
  ncu -i /tmp/prof_items.ncu-rep --page raw --csv > gpurun_out/prof_items_raw.csv 2> gpurun_out/prof_items_raw.err
  for k in gemm_bf16_cg2 attention_tc attention_kernel score_tile; do
    ncu -i /tmp/prof_users.ncu-rep --page source --csv -k regex:$k -c 1 > gpurun_out/src_users_$k.csv 2>> gpurun_out/prof_users_raw.err
  done
  ls -la /tmp/*.ncu-rep >> gpurun_out/prof_users_raw.err
  ;;
kernels)
  timeout 300 python tools/gpu_bench_kernels.py attention rowwise scoring joint train > gpurun_out/kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/kernels.log
  ;;
esac
done
ls -la gpurun_out
