#!/bin/bash
# Round 2, call V (8 GPUs): the bench at N = 8 exactly as the driver launches it (users / items weak scaling, training strong
# scaling with the gradient all-reduces inside the CUDA graph).
set -u
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "rc=$?" >> gpurun_out/bench_n8.err
tail -2 gpurun_out/bench_n8.err
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/bench_n8.json") if l.startswith("{")][0])
t = d["train"]
print(f"N=8: users/s {d['value']:.0f} ms {d['ms_per_step']:.2f} e2e {d['e2e']['value']:.0f} items/s {d['items']['value']:.0f} e2e {d['items']['e2e']['value']:.0f} parity {d['parity_vs_gpu']['ok']} train items/s {t['value']:.0f} ms {t['ms_per_step']:.2f} {t['mode']} variants {t['cuda_graph'].get('variants_ms_per_step')} eager {t['eager']['ms_per_step']:.2f}")
P
