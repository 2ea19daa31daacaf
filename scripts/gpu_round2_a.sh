#!/bin/bash
# first GPU call of round 2: new tests, eval A/B timing, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_regroup.py tests/test_gpu_eval.py -x -q -m gpu > gpurun_out/r2a_tests_eval.log 2>&1; echo "eval tests rc=$?" 
timeout 600 python scripts/eval_ab.py 1e9 5 > gpurun_out/r2a_eval_ab.log 2>&1; echo "eval_ab rc=$?"
timeout 1200 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_regroup.py --deselect tests/test_gpu_eval.py > gpurun_out/r2a_tests_rest.log 2>&1; echo "rest tests rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r2a_tests_eval.log; cat gpurun_out/r2a_eval_ab.log; tail -5 gpurun_out/r2a_tests_rest.log
