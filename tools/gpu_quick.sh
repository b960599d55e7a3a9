#!/bin/bash
# Quick GPU call: selected tests + one bench run without the CPU / eager context legs.
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_modules_gpu.py tests/test_ranking_gpu.py -m gpu -q --timeout 200 -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_quick.log
timeout 400 python bench.py --no-cpu-baseline --train-batch 0 > gpurun_out/b_quick.json 2> gpurun_out/b_quick.err; echo "rc=$?" >> gpurun_out/b_quick.err
