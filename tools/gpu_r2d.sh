#!/bin/bash
# Round 2, call D (1 GPU): GPU tests with the CTA-pair scoring kernel, bench (user + item stages), per-shape scoring timings,
# sanitizer pass over every kernel family (memcheck) and the pair scoring kernel (racecheck / synccheck).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --durations=8 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --train-batch 0 --no-cpu-baseline --steps 4 > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err; echo "rc=$?" >> gpurun_out/bench_plain.err
timeout 300 python tools/gpu_score_shapes.py > gpurun_out/score_shapes.log 2>&1; echo "rc=$?" >> gpurun_out/score_shapes.log
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --print-limit 20 --error-exitcode 9 python tools/sanitizer_cases.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_memcheck.log
timeout 600 $CS --tool racecheck --print-limit 20 --error-exitcode 9 python tools/sanitizer_cases.py score_topk > gpurun_out/sanitizer_racecheck_score_topk.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_racecheck_score_topk.log
timeout 600 $CS --tool synccheck --print-limit 20 --error-exitcode 9 python tools/sanitizer_cases.py score_topk train_kernels rowwise lists > gpurun_out/sanitizer_synccheck2.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_synccheck2.log
tail -n 3 gpurun_out/sanitizer_*.log
