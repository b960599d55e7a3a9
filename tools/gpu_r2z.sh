#!/bin/bash
# Round 2, call Z (1 GPU): score_topk with pool-dependent users per pass: ranking parity tests (incl. the per-rank shapes of the
# 4- and 8-GPU runs), shape timings.
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_ranking_gpu.py tests/test_bench_shapes_gpu.py -m gpu -q --timeout 200 -p no:cacheprovider -x > gpurun_out/pytest_ranking.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ranking.log
tail -4 gpurun_out/pytest_ranking.log
timeout 200 python tools/gpu_score_shapes.py > gpurun_out/score_shapes.log 2>&1; echo "rc=$?" >> gpurun_out/score_shapes.log
cat gpurun_out/score_shapes.log
