#!/bin/bash
# compute-sanitizer over one small launch of every kernel family (tools/sanitizer_cases.py): memcheck on everything,
# racecheck + synccheck on the shared-memory / mbarrier / TMEM kernels.  Logs -> gpurun_out/sanitizer_*.log
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 120 python tools/sanitizer_cases.py > gpurun_out/sanitizer_plain.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_plain.log
timeout 900 $CS --tool memcheck --print-limit 20 --error-exitcode 9 python tools/sanitizer_cases.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_memcheck.log
for c in gemm_cg2 attention_tc kv_attention score_topk attention_small; do
  timeout 600 $CS --tool racecheck --print-limit 20 --error-exitcode 9 python tools/sanitizer_cases.py $c > gpurun_out/sanitizer_racecheck_$c.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_racecheck_$c.log
done
timeout 600 $CS --tool synccheck --print-limit 20 --error-exitcode 9 python tools/sanitizer_cases.py gemm_cg2 attention_tc kv_attention score_topk > gpurun_out/sanitizer_synccheck.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_synccheck.log
tail -n 4 gpurun_out/sanitizer_*.log
