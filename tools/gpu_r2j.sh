#!/bin/bash
# Round 2, call J (1 GPU): reverse traversal of the streaming kernels (LayerNorm, small-tile attention) - tests, then an A/B
# of the bench on one box (UNIREC_STREAM_REVERSE = 1 (alternating traversal directions) / 0 (all forward), twice each, interleaved).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py tests/test_train_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_reverse.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_reverse.log
tail -3 gpurun_out/pytest_reverse.log
for rep in 1 2; do
  for rev in 1 0; do
    UNIREC_STREAM_REVERSE=$rev timeout 300 python bench.py --train-batch 0 --no-cpu-baseline --steps 4 > gpurun_out/bench_rev${rev}_$rep.json 2> gpurun_out/bench_rev${rev}_$rep.err
  done
done
python - <<'P'
import json
for rep in (1, 2):
    for rev in (1, 0):
        d = json.loads([l for l in open(f"gpurun_out/bench_rev{rev}_{rep}.json") if l.startswith("{")][0])
        print(f"reverse={rev} rep={rep}: users/s {d['value']:.0f} (ms {d['ms_per_step']:.2f}, clock {d['clocks']['sm_mhz']}) items/s {d['items']['value']:.0f}")
P
