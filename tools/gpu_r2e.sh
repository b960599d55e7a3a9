#!/bin/bash
# Round 2, call E (2 GPUs): the 2-rank NCCL tests (data-parallel gradients eager + captured in the graph, sharded ranking
# with both list exchanges), then the bench at N = 2 (all-to-all exchange, training with the collectives inside the graph).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > gpurun_out/smi2.txt 2>&1
timeout 60 tools/bin/probe_mixed_cta_group > gpurun_out/probe_mixed_cta_group.log 2>&1; echo "rc=$?" >> gpurun_out/probe_mixed_cta_group.log
timeout 900 python -m pytest tests/test_dp_gpu.py -m gpu -q --timeout 400 -p no:cacheprovider -s > gpurun_out/pytest_dp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_dp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?" >> gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --train-only --train-wire fp32 > gpurun_out/bench_n2_train_fp32wire.json 2> gpurun_out/bench_n2_train_fp32wire.err; echo "rc=$?" >> gpurun_out/bench_n2_train_fp32wire.err
tail -n 5 gpurun_out/pytest_dp.log
