#!/bin/bash
# Round 2, call N (1 GPU): ncu --set full + source-level sampling of ONE attention_tc launch (512 users x 64 x 1600, ragged
# mask) for a per-phase breakdown of the softmax warps' tile period.
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 5 -c 1 -f -o /tmp/prof_attn \
    python tools/gpu_attn_tc_time.py > gpurun_out/prof_attn.out 2>&1
ncu -i /tmp/prof_attn.ncu-rep --page raw --csv > gpurun_out/prof_attn_raw.csv 2> gpurun_out/prof_attn_raw.err
ncu -i /tmp/prof_attn.ncu-rep --page source --csv > gpurun_out/src_attn_tc.csv 2>> gpurun_out/prof_attn_raw.err
tail -5 gpurun_out/prof_attn.out; ls -la gpurun_out/src_attn_tc.csv
