#!/bin/bash
# Round 2, call K (4 GPUs): the bench at N = 4 exactly as the driver launches it (all-to-all list exchange, training block with
# the gradient all-reduces captured inside the CUDA graph, bf16 wire).
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "rc=$?" >> gpurun_out/bench_n4.err
tail -3 gpurun_out/bench_n4.err
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/bench_n4.json") if l.startswith("{")][0])
print("users/s", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "items/s", round(d["items"]["value"]))
print("parity", d["parity_vs_gpu"]["ok"], d["parity_vs_gpu"]["picks_outside_tie_tolerance"])
t = d["train"]
print("train items/s", round(t["value"]), "ms", round(t["ms_per_step"], 2), t["mode"], t["cuda_graph"].get("variants_ms_per_step"), "eager", round(t["eager"]["ms_per_step"], 2), "bytes", t["allreduce_bytes_per_step"])
P
