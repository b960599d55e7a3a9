#!/bin/bash
# Training-path GPU call: train tests, per-kernel timings of the training kernels, the cfg-2 block of the bench.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_train_gpu.py -m gpu -q --timeout 120 -p no:cacheprovider > gpurun_out/pytest_train.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_train.log
timeout 200 python tools/gpu_bench_kernels.py train > gpurun_out/kernels_train.log 2>&1
timeout 300 python bench.py --train-only > gpurun_out/b_train.json 2> gpurun_out/b_train.err; echo "rc=$?" >> gpurun_out/b_train.err
