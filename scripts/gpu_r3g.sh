#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_regroup.py tests/test_gpu_eval.py -x -q -m gpu > gpurun_out/r3g_tests_eval.log 2>&1; echo "eval tests rc=$?"
tail -5 gpurun_out/r3g_tests_eval.log
timeout 600 python scripts/eval_sweep.py > gpurun_out/r3g_eval_sweep.md 2> gpurun_out/r3g_eval_sweep.err; echo "sweep rc=$?"; grep "real32" gpurun_out/r3g_eval_sweep.md | grep "1e+09\|splde"
