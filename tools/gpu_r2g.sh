#!/bin/bash
# Round 2, call G (1 GPU): UMMA variant of the fused kernel after the polling fix - tests, bench, ncu.
set -u
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_kv_attention_gpu.py -m gpu -q --timeout 60 -x -p no:cacheprovider -k "umma" > gpurun_out/pytest_kvattn_umma.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_kvattn_umma.log
if grep -q "pytest rc=0" gpurun_out/pytest_kvattn_umma.log; then
  UNIREC_KV_ATTENTION_IMPL=umma timeout 300 python bench.py --train-batch 0 --no-cpu-baseline --steps 4 --fused-kv 1 > gpurun_out/bench_fusedkv_umma.json 2> gpurun_out/bench_fusedkv_umma.err; echo "rc=$?" >> gpurun_out/bench_fusedkv_umma.err
  UNIREC_KV_ATTENTION_IMPL=umma timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'kv_attention_umma' -c 1 -f -o gpurun_out/prof_kvumma \
      python bench.py --steps 1 --warmup 1 --pool-items 131072 --users-per-gpu 512 --no-cpu-baseline --train-batch 0 --fused-kv 1 --profile-range users > gpurun_out/prof_kvumma.out 2>&1
  ncu -i gpurun_out/prof_kvumma.ncu-rep --page raw --csv > gpurun_out/prof_kvumma_raw.csv 2> gpurun_out/prof_kvumma_raw.err
  ncu -i gpurun_out/prof_kvumma.ncu-rep --page source --csv -k regex:kv_attention_umma -c 1 > gpurun_out/src_kvumma.csv 2>> gpurun_out/prof_kvumma_raw.err
  rm -f gpurun_out/prof_kvumma.ncu-rep
fi
timeout 300 python bench.py --train-batch 0 --no-cpu-baseline --steps 4 > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err; echo "rc=$?" >> gpurun_out/bench_plain.err
ls -la gpurun_out | tail -8
