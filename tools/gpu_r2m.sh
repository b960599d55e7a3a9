#!/bin/bash
# Round 2, call M (1 GPU): (1) the full GPU test suite with the LayerNorm fold on by default (deterministic per-tile
# statistics), (2) attention_tc with 0 / 25 / 50 / 75 % of the softmax exponentials on the FMA pipe: parity + timing alone
# and inside the step's kernel mix.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for emu in 0 1 2 3; do
  UNIREC_ATTN_EXP_EMU=$emu timeout 300 python tools/gpu_attn_tc_time.py > gpurun_out/attn_tc_emu$emu.log 2>&1; echo "rc=$?" >> gpurun_out/attn_tc_emu$emu.log
  cat gpurun_out/attn_tc_emu$emu.log | tail -8
done
