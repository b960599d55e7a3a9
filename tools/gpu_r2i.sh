#!/bin/bash
# Round 2, call I (2 GPUs): the 2-rank NCCL tests after the fixes (no NCCL teardown in the workers, one-rank group for the
# single-GPU ranker); short timeouts.
set -u
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_dp_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -s > gpurun_out/pytest_dp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_dp.log
grep -E "rank |passed|failed|rc=" gpurun_out/pytest_dp.log | tail -20
