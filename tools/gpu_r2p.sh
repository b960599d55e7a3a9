#!/bin/bash
# Round 2, call P (1 GPU): CTA-pair GEMM with 16 epilogue warps per CTA (UNIREC_GEMM_EPI_WARPS=16): kernel + module parity
# tests, epilogue cost under sustained load next to the 8-warp kernel, interleaved bench A/B.
set -u
mkdir -p gpurun_out
UNIREC_GEMM_EPI_WARPS=16 timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py tests/test_bench_shapes_gpu.py -m gpu -q --timeout 120 -p no:cacheprovider -x -k "linear or qformer or gemm or 4096" > gpurun_out/pytest_epi16.log 2>&1
rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_epi16.log
tail -5 gpurun_out/pytest_epi16.log
if [ $rc -ne 0 ]; then exit 0; fi
for ew in 16 8; do
  UNIREC_GEMM_EPI_WARPS=$ew timeout 200 python tools/gpu_gemm_modes.py > gpurun_out/gemm_modes_epi$ew.log 2>&1
  echo "--- epilogue warps $ew"; cat gpurun_out/gemm_modes_epi$ew.log
done
for rep in 1 2; do
  for ew in 16 8; do
    UNIREC_GEMM_EPI_WARPS=$ew timeout 300 python bench.py --train-batch 0 --no-cpu-baseline --steps 4 > gpurun_out/bench_epi${ew}_$rep.json 2> gpurun_out/bench_epi${ew}_$rep.err
  done
done
python - <<'P'
import json
for rep in (1, 2):
    for ew in (16, 8):
        try:
            d = json.loads([l for l in open(f"gpurun_out/bench_epi{ew}_{rep}.json") if l.startswith("{")][0])
            print(f"epi_warps={ew} rep={rep}: users/s {d['value']:.0f} (ms {d['ms_per_step']:.2f}, clock {d['clocks']['sm_mhz']}) items/s {d['items']['value']:.0f}")
        except Exception as e:
            print("epi", ew, rep, "failed", e)
P
