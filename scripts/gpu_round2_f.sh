#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2f_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2f_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
for k in ('metric','value','evals_per_s','fit_ms','eval_ms','ms_per_step','stages_ms','roofline','cpu_baseline','cpu_baseline_algorithm_matched','e2e','e2e_pageable','gpu_launches','clocks'): print(k, json.dumps(b.get(k))[:400])
r=json.loads(open('gpurun_out/r2f_bench_ref.json').read().strip().splitlines()[-1]); print('ref', r['metric'], r['value'])
PY
