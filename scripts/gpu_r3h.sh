#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fit.py tests/test_gpu_scale.py tests/test_gpu_ortho.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r3h_tests_fit.log 2>&1; echo "fit tests rc=$?"
tail -15 gpurun_out/r3h_tests_fit.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r3h_bench.json 2> gpurun_out/r3h_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r3h_bench.json')); print(d['fit_ms'], d['eval_ms'], d['stages_ms'])"
