#!/bin/bash
# Round 2, call B: GPU tests after the tolerance fixes, A/B of the fused K/V-projection + attention path against the
# materialised path on one box, launch list + ncu --set full of the fused kernel (512 users = 128 groups of 4 users).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python -m pytest tests/test_kv_attention_gpu.py -m gpu -q --timeout 120 -p no:cacheprovider -s > gpurun_out/pytest_kvattn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_kvattn.log
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --durations=10 -p no:cacheprovider --deselect tests/test_kv_attention_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --train-batch 0 --no-cpu-baseline --steps 4 > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err; echo "rc=$?" >> gpurun_out/bench_plain.err
timeout 600 python bench.py --train-batch 0 --no-cpu-baseline --steps 4 --fused-kv 1 > gpurun_out/bench_fusedkv.json 2> gpurun_out/bench_fusedkv.err; echo "rc=$?" >> gpurun_out/bench_fusedkv.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_users_fusedkv.csv \
    python bench.py --steps 1 --warmup 1 --pool-items 131072 --users-per-gpu 512 --no-cpu-baseline --train-batch 0 --fused-kv 1 --profile-range users > gpurun_out/launches_users_fusedkv.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'kv_attention' -c 2 -f -o gpurun_out/prof_kvattn \
    python bench.py --steps 1 --warmup 1 --pool-items 131072 --users-per-gpu 512 --no-cpu-baseline --train-batch 0 --fused-kv 1 --profile-range users > gpurun_out/prof_kvattn.out 2>&1
ncu -i gpurun_out/prof_kvattn.ncu-rep --page raw --csv > gpurun_out/prof_kvattn_raw.csv 2> gpurun_out/prof_kvattn_raw.err
ncu -i gpurun_out/prof_kvattn.ncu-rep --page source --csv -k regex:kv_attention_fused -c 1 > gpurun_out/src_kvattn.csv 2>> gpurun_out/prof_kvattn_raw.err
ls -la gpurun_out
