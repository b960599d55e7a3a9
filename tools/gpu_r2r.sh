#!/bin/bash
# Round 2, call R (1 GPU): final code - GPU tests, smoke, full bench line, reference arm, ncu launch lists of one user step /
# two item chunks, ncu --set full of the first CTA-pair GEMM launches of a user chunk (the dominant kernel + a folded GEMM).
set -u
mkdir -p gpurun_out
bash tools/gpu_round.sh tests bench launches
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo "rc=$?" >> gpurun_out/bench_reference_arm.err
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:gemm_bf16_cg2 -c 4 -f -o /tmp/prof_gemm \
    python bench.py --steps 1 --warmup 1 --pool-items 131072 --users-per-gpu 512 --no-cpu-baseline --train-batch 0 --profile-range users \
    > gpurun_out/prof_gemm.out 2>&1
ncu -i /tmp/prof_gemm.ncu-rep --page raw --csv > gpurun_out/prof_gemm_raw.csv 2> gpurun_out/prof_gemm_raw.err
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -2 gpurun_out/bench_full.err; tail -2 gpurun_out/bench_reference_arm.err
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/bench_full.json") if l.startswith("{")][0])
print("users/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), "items/s", round(d["items"]["value"]), "e2e", round(d["items"]["e2e"]["value"]), "train", round(d["train"]["value"]), "roofline", round(d["roofline"]["frac"], 3), "parity", d["parity_vs_gpu"]["ok"], "cpu", d["cpu_baseline"])
P
