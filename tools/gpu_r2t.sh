#!/bin/bash
# Round 2, call T (1 GPU): compute-sanitizer over the kernels added or changed late in round 2 - the LayerNorm-folding GEMM
# epilogue (gemm_ln), attention_tc_kernel with P in tensor memory (default) and with P in shared memory, the two-group
# attention kernel (UNIREC_ATTENTION_PP=1): memcheck, racecheck, synccheck.  Then the attention timing of the final code.
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # name, env, tool, cases...
  local name=$1 envs=$2 tool=$3; shift 3
  env $envs timeout 600 $CS --tool $tool --print-limit 20 --error-exitcode 9 python tools/sanitizer_cases.py "$@" > gpurun_out/sanitizer_$name.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_$name.log
  echo "== $name: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|rc=' gpurun_out/sanitizer_$name.log | tr '\n' ' ')"
}
run memcheck_ln_tc "A=1" memcheck gemm_ln gemm_cg2 attention_tc
run memcheck_tc_psmem "UNIREC_ATTENTION_PTMEM=0" memcheck attention_tc
run memcheck_pp "UNIREC_ATTENTION_PP=1" memcheck attention_tc
run racecheck_ln "A=1" racecheck gemm_ln
run racecheck_tc_ptmem "A=1" racecheck attention_tc
run racecheck_pp "UNIREC_ATTENTION_PP=1" racecheck attention_tc
run synccheck_all "A=1" synccheck gemm_ln attention_tc
run synccheck_pp "UNIREC_ATTENTION_PP=1" synccheck attention_tc
for pt in 1 0; do
  UNIREC_ATTENTION_PTMEM=$pt timeout 120 python tools/gpu_attn_tc_time.py > gpurun_out/attn_ptmem$pt.log 2>&1; echo "rc=$?" >> gpurun_out/attn_ptmem$pt.log
  tail -4 gpurun_out/attn_ptmem$pt.log
done
