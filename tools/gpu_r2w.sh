#!/bin/bash
# Round 2, call W (1 GPU): the full bench line of the final code (with the kernels_alone block); a short small-pool run first.
set -u
mkdir -p gpurun_out
timeout 200 python bench.py --pool-items 131072 --users-per-gpu 512 --steps 1 --warmup 1 --no-cpu-baseline --train-batch 0 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/bench_small.json") if l.startswith("{")][0])
print("small run kernels_alone:", {k: (round(v["frac"], 3) if isinstance(v, dict) else v) for k, v in (d.get("kernels_alone") or {}).items()})
P
timeout 700 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?" >> gpurun_out/bench_full.err
tail -2 gpurun_out/bench_full.err
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/bench_full.json") if l.startswith("{")][0])
print("users/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "items/s", round(d["items"]["value"]), "train", round(d["train"]["value"]), "clock", d["clocks"]["sm_mhz"])
for k, v in (d.get("kernels_alone") or {}).items():
    print(" ", k, v if not isinstance(v, dict) else {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
P
