#!/bin/bash
# 8 GPUs: the contract bench at N = 8 (weak line + strong sub-record + e2e) of the final build
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r5r_bench_n8.json 2> gpurun_out/r5r_bench_n8.err; echo "bench n8 rc=$?"; tail -3 gpurun_out/r5r_bench_n8.err
python - <<'PY'
import json
j = json.loads([l for l in open("gpurun_out/r5r_bench_n8.json") if l.startswith("{")][-1])
print({k: j[k] for k in ("value", "evals_per_s", "ms_per_step", "fit_ms", "eval_ms")})
print("strong", j.get("strong"))
PY
