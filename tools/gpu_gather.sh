#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_ranking_gpu.py tests/test_joint_gpu.py -m gpu -q --timeout 100 -p no:cacheprovider -k "gather or more_users" > gpurun_out/pytest_gather.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gather.log
if grep -q "rc=0" gpurun_out/pytest_gather.log; then
timeout 300 python bench.py --no-cpu-baseline --train-batch 0 --fused-gather 1 > gpurun_out/b_gather.json 2> gpurun_out/b_gather.err; echo "rc=$?" >> gpurun_out/b_gather.err
timeout 300 python bench.py --no-cpu-baseline --train-batch 0 --fused-gather 0 > gpurun_out/b_plain.json 2> gpurun_out/b_plain.err; echo "rc=$?" >> gpurun_out/b_plain.err
fi
