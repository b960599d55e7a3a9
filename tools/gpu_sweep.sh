#!/bin/bash
# One gpurun call: joint/train tests, joint kernel timings, bench sweeps over the user-chunk and item-batch sizes.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_joint_gpu.py tests/test_train_gpu.py -m gpu -q --timeout 120 -p no:cacheprovider > gpurun_out/pytest_joint.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_joint.log
timeout 200 python tools/gpu_bench_kernels.py joint > gpurun_out/kernels_joint.log 2>&1
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/b_base.json 2> gpurun_out/b_base.err
timeout 300 python bench.py --no-cpu-baseline --train-batch 0 --kv-gb 28 > gpurun_out/b_kv28.json 2> gpurun_out/b_kv28.err
timeout 300 python bench.py --no-cpu-baseline --train-batch 0 --kv-gb 56 --item-batch 8192 > gpurun_out/b_kv56_ib8192.json 2> gpurun_out/b_kv56.err
ls -la gpurun_out
