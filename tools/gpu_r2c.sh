#!/bin/bash
# Round 2, call C: GPU tests (incl. the full-size training parity test), A/B of the layer-major user encoder against the
# chunk-major loop of round 1 on one box, compute-sanitizer pass, ncu --set full of the K/V projection at its new shape.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --durations=12 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --train-batch 0 --no-cpu-baseline --steps 4 > gpurun_out/bench_layer_major.json 2> gpurun_out/bench_layer_major.err; echo "rc=$?" >> gpurun_out/bench_layer_major.err
timeout 600 python bench.py --train-batch 0 --no-cpu-baseline --steps 4 --users-per-call 512 > gpurun_out/bench_chunk_major.json 2> gpurun_out/bench_chunk_major.err; echo "rc=$?" >> gpurun_out/bench_chunk_major.err
bash tools/gpu_sanitizer.sh > gpurun_out/sanitizer_tail.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'gemm_bf16_cg2' -c 1 -f -o gpurun_out/prof_kvgemm \
    python bench.py --steps 1 --warmup 1 --pool-items 131072 --no-cpu-baseline --train-batch 0 --profile-range users > gpurun_out/prof_kvgemm.out 2>&1
ncu -i gpurun_out/prof_kvgemm.ncu-rep --page raw --csv > gpurun_out/prof_kvgemm_raw.csv 2> gpurun_out/prof_kvgemm_raw.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_users.csv \
    python bench.py --steps 1 --warmup 1 --pool-items 131072 --no-cpu-baseline --train-batch 0 --profile-range users > gpurun_out/launches_users.out 2>&1
rm -f gpurun_out/prof_kvgemm.ncu-rep
ls -la gpurun_out
