#!/bin/bash
# Round 2, call Q (1 GPU): the cfg-2 training step at the per-GPU batch sizes of the 1 / 4 / 8 GPU runs (1024 / 256 / 128
# items), no collectives: how far the compute of a rank shrinks with its batch; per-kernel launch list at 128 items.
set -u
mkdir -p gpurun_out
for b in 1024 256 128; do
  timeout 300 python bench.py --train-only --train-batch $b > gpurun_out/train_b$b.json 2> gpurun_out/train_b$b.err; echo "rc=$?" >> gpurun_out/train_b$b.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train_b128.csv python bench.py --train-only --train-batch 128 --train-steps 1 --profile-range train > gpurun_out/launches_train_b128.out 2>&1
python - <<'P'
import json
for b in (1024, 256, 128):
    try:
        t = json.loads([l for l in open(f"gpurun_out/train_b{b}.json") if l.startswith("{")][0])["train"]
        print(b, "items: graph", round(t["cuda_graph"]["ms_per_step"], 3), "ms  eager", round(t["eager"]["ms_per_step"], 3), "ms  host enqueue", round(t["eager"]["host_enqueue_ms_per_step"], 2), "launches", t["gpu_launches_per_step"])
    except Exception as e:
        print(b, "failed", e)
P
