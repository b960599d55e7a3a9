#!/bin/bash
# Round 2, call O (1 GPU): the two-group ("ping-pong") long-key attention kernel (attention_pp.cu): the attention parity
# tests with UNIREC_ATTENTION_PP=1, then timing alone / inside the step's kernel mix next to attention_tc_kernel.
set -u
mkdir -p gpurun_out
UNIREC_ATTENTION_PP=1 timeout 150 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 60 -p no:cacheprovider -x -k "attention" > gpurun_out/pytest_attn_pp.log 2>&1
rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_attn_pp.log
tail -15 gpurun_out/pytest_attn_pp.log
if [ $rc -ne 0 ]; then exit 0; fi
UNIREC_ATTENTION_PP=1 timeout 150 python -m pytest tests/test_modules_gpu.py -m gpu -q --timeout 60 -p no:cacheprovider -x -k "user_qformer" >> gpurun_out/pytest_attn_pp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_attn_pp.log
tail -4 gpurun_out/pytest_attn_pp.log
for pp in 1 0; do
  UNIREC_ATTENTION_PP=$pp timeout 120 python tools/gpu_attn_tc_time.py > gpurun_out/attn_pp$pp.log 2>&1; echo "rc=$?" >> gpurun_out/attn_pp$pp.log
  tail -8 gpurun_out/attn_pp$pp.log
done
