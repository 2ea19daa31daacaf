#!/bin/bash
mkdir -p gpurun_out
python scripts/dbg_regroup.py 2>&1 | grep -v "^\[\[\| \[" | tail -20
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2b_tests.log
timeout 600 python scripts/eval_ab.py 1e9 4 3 > gpurun_out/r2b_eval_ab.log 2>&1; cat gpurun_out/r2b_eval_ab.log
