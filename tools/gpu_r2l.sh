#!/bin/bash
# Round 2, call L (1 GPU): LayerNorm folded into the GEMMs around it (unirec_linear_ln_bf16) - kernel + module parity tests,
# then an interleaved A/B of the bench on one box (--fold-ln 1 / 0, twice each).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -s -k "folded or linear_ln" > gpurun_out/pytest_foldln.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_foldln.log
grep -E "folded|launches|passed|failed|rc=|Error|error" gpurun_out/pytest_foldln.log | tail -40
for rep in 1 2; do
  for f in 1 0; do
    timeout 300 python bench.py --train-batch 0 --no-cpu-baseline --steps 4 --fold-ln $f > gpurun_out/bench_foldln${f}_$rep.json 2> gpurun_out/bench_foldln${f}_$rep.err
  done
done
python - <<'P'
import json
for rep in (1, 2):
    for f in (1, 0):
        try:
            d = json.loads([l for l in open(f"gpurun_out/bench_foldln{f}_{rep}.json") if l.startswith("{")][0])
            print(f"fold_ln={f} rep={rep}: users/s {d['value']:.0f} (ms {d['ms_per_step']:.2f}, clock {d['clocks']['sm_mhz']}) items/s {d['items']['value']:.0f} parity {d['parity_vs_gpu']['ok']}")
        except Exception as e:
            print("fold_ln", f, rep, "failed", e)
P
tail -5 gpurun_out/bench_foldln1_1.err
