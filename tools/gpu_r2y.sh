#!/bin/bash
# Round 2, call Y (1 GPU): final code (attention_tc_kernel<PTMEM, QTMA> default) - sanitizer on the attention kernel, the full
# GPU test suite, smoke, the full bench line.
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 600 $CS --tool $tool --print-limit 20 --error-exitcode 9 python tools/sanitizer_cases.py attention_tc > gpurun_out/sanitizer_${tool}_tc_qtma.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_${tool}_tc_qtma.log
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|rc=' gpurun_out/sanitizer_${tool}_tc_qtma.log | tr '\n' ' ')"
done
bash tools/gpu_round.sh tests bench > /dev/null 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log; tail -1 gpurun_out/bench_full.err
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/bench_full.json") if l.startswith("{")][0])
print("users/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), "items/s", round(d["items"]["value"]), "e2e", round(d["items"]["e2e"]["value"]), "train", round(d["train"]["value"]), "clock", d["clocks"]["sm_mhz"], "parity", d["parity_vs_gpu"]["ok"], d["parity_vs_gpu"]["topk_overlap"])
o = d["roofline"]["other_kernels"]
print({k: (round(v["frac"], 3), round(v["avg_launch_ms"], 3)) for k, v in o["attention_by_shape"].items()}, "score", round(o["score_topk"]["frac"], 3))
print({k: round(v["frac"], 3) for k, v in d["kernels_alone"].items()})
P
