#!/bin/bash
# Round 2, call X (1 GPU): attention_tc_kernel with the stacked queries loaded by TMA, lane-masked S MMAs and three V stages
# (UNIREC_ATTENTION_QTMA=1): attention parity tests, user Q-Former goldens, timing next to the Q2-by-softmax-warps kernel.
set -u
mkdir -p gpurun_out
UNIREC_ATTENTION_QTMA=1 timeout 150 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 60 -p no:cacheprovider -x -k "attention and not two_group" > gpurun_out/pytest_attn_qtma.log 2>&1
rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_attn_qtma.log
tail -25 gpurun_out/pytest_attn_qtma.log
if [ $rc -ne 0 ]; then exit 0; fi
UNIREC_ATTENTION_QTMA=1 timeout 150 python -m pytest tests/test_modules_gpu.py -m gpu -q --timeout 60 -p no:cacheprovider -x -k "user_qformer" >> gpurun_out/pytest_attn_qtma.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_attn_qtma.log
tail -4 gpurun_out/pytest_attn_qtma.log
for qt in 1 0; do
  UNIREC_ATTENTION_QTMA=$qt timeout 120 python tools/gpu_attn_tc_time.py > gpurun_out/attn_qtma$qt.log 2>&1; echo "rc=$?" >> gpurun_out/attn_qtma$qt.log
  tail -4 gpurun_out/attn_qtma$qt.log
done
