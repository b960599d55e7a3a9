#!/bin/bash
# Round 2, call U (4 GPUs): final code - the 2-rank NCCL tests (data-parallel gradients, sharded ranking), then the bench at
# N = 4 and N = 2 exactly as the driver launches it.
set -u
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_dp_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -s > gpurun_out/pytest_dp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_dp.log
grep -E "passed|failed|rc=" gpurun_out/pytest_dp.log | tail -5
for n in 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "rc=$?" >> gpurun_out/bench_n$n.err
  tail -2 gpurun_out/bench_n$n.err
done
python - <<'P'
import json
for n in (4, 2):
    try:
        d = json.loads([l for l in open(f"gpurun_out/bench_n{n}.json") if l.startswith("{")][0])
        t = d["train"]
        print(f"N={n}: users/s {d['value']:.0f} ms {d['ms_per_step']:.2f} e2e {d['e2e']['value']:.0f} items/s {d['items']['value']:.0f} parity {d['parity_vs_gpu']['ok']} train items/s {t['value']:.0f} ms {t['ms_per_step']:.2f} {t['mode']}")
    except Exception as e:
        print(n, "failed", e)
P
