#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_eval.py tests/test_gpu_fit.py -x -q -m gpu 2>&1 | tail -15
timeout 600 python scripts/eval_sweep.py 1e9 > gpurun_out/r2_eval_sweep.md 2>&1; tail -60 gpurun_out/r2_eval_sweep.md
