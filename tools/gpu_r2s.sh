#!/bin/bash
# Round 2, call S (1 GPU): attention_tc_kernel with P as a tensor-memory operand of the second MMA (UNIREC_ATTENTION_PTMEM=1):
# attention parity tests, user Q-Former goldens, timing alone / in the step's kernel mix next to the shared-memory-P kernel.
set -u
mkdir -p gpurun_out
UNIREC_ATTENTION_PTMEM=1 timeout 150 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 60 -p no:cacheprovider -x -k "attention and not two_group" > gpurun_out/pytest_attn_ptmem.log 2>&1
rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_attn_ptmem.log
tail -25 gpurun_out/pytest_attn_ptmem.log
if [ $rc -ne 0 ]; then exit 0; fi
UNIREC_ATTENTION_PTMEM=1 timeout 150 python -m pytest tests/test_modules_gpu.py -m gpu -q --timeout 60 -p no:cacheprovider -x -k "user_qformer" >> gpurun_out/pytest_attn_ptmem.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_attn_ptmem.log
tail -4 gpurun_out/pytest_attn_ptmem.log
for pt in 1 0; do
  UNIREC_ATTENTION_PTMEM=$pt timeout 120 python tools/gpu_attn_tc_time.py > gpurun_out/attn_ptmem$pt.log 2>&1; echo "rc=$?" >> gpurun_out/attn_ptmem$pt.log
  tail -8 gpurun_out/attn_ptmem$pt.log
done
