#!/bin/bash
# 2-GPU call: joint tests + kernel timings on GPU 0, then the bench under torchrun (N = 2) and its reference arm.
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_joint_gpu.py -m gpu -q --timeout 120 -p no:cacheprovider > gpurun_out/pytest_joint.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_joint.log
timeout 150 python tools/gpu_bench_kernels.py joint > gpurun_out/kernels_joint.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?" >> gpurun_out/bench_n2.err
ls -la gpurun_out
