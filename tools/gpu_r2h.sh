#!/bin/bash
# Round 2, call H (1 GPU): where does the tile period of the UMMA fused kernel go?  Timing-only variants (results are wrong
# with the MMAs skipped): no S MMAs, no PV MMAs, neither.
set -u
mkdir -p gpurun_out
for dbg in 0 1 2 3; do
  UNIREC_KV_DEBUG=$dbg UNIREC_KV_ATTENTION_IMPL=umma timeout 200 python tools/gpu_kvattn_time.py > gpurun_out/kvattn_time_dbg$dbg.log 2>&1
done
UNIREC_KV_ATTENTION_IMPL=mma_sync timeout 200 python tools/gpu_kvattn_time.py > gpurun_out/kvattn_time_mma_sync.log 2>&1
cat gpurun_out/kvattn_time_*.log
