#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_regroup.py tests/test_gpu_eval.py -x -q -m gpu > gpurun_out/r3d_tests_eval.log 2>&1; echo "eval tests rc=$?"
tail -5 gpurun_out/r3d_tests_eval.log
timeout 600 python scripts/eval_ab.py 1e9 3 1,2,3 2>&1 | tee gpurun_out/r3d_eval_ab.log
timeout 600 python scripts/eval_sweep.py > gpurun_out/r3d_eval_sweep.md 2> gpurun_out/r3d_eval_sweep.err; echo "sweep rc=$?"; tail -40 gpurun_out/r3d_eval_sweep.md
