#!/bin/bash
# Round 2, call A: GPU tests (incl. the new parity-at-bench-shape, encoder, fused-kv tests), smoke, bench A/B of the fused
# K/V-projection + attention path against the materialised path on the same box, reference arm.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
# the fused kernel first, alone and under a short timeout: a hang must not take the whole call with it
timeout 300 python -m pytest tests/test_kv_attention_gpu.py -m gpu -q --timeout 120 -p no:cacheprovider -s > gpurun_out/pytest_kvattn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_kvattn.log
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --durations=15 -p no:cacheprovider --deselect tests/test_kv_attention_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --train-batch 0 --no-cpu-baseline --steps 4 > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err; echo "rc=$?" >> gpurun_out/bench_plain.err
if grep -q "pytest rc=0" gpurun_out/pytest_kvattn.log; then
  timeout 600 python bench.py --train-batch 0 --no-cpu-baseline --steps 4 --fused-kv 1 > gpurun_out/bench_fusedkv.json 2> gpurun_out/bench_fusedkv.err; echo "rc=$?" >> gpurun_out/bench_fusedkv.err
fi
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?" >> gpurun_out/bench_full.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err
ls -la gpurun_out
