#!/bin/bash
# round 2, session 2: uniform-form evaluation -- tests, then A/B timing exact vs uniform
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_regroup.py tests/test_gpu_eval.py -x -q -m gpu > gpurun_out/r3a_tests_eval.log 2>&1; echo "eval tests rc=$?"
tail -15 gpurun_out/r3a_tests_eval.log
for b in exact uniform; do
  echo "== basis $b"
  SPLPAK_B200_BASIS=$b timeout 600 python scripts/eval_ab.py 1e9 4 3,2 2>&1 | tee -a gpurun_out/r3a_eval_ab_$b.log
done
for w in 16 32; do
  echo "== uniform, $w warps"
  SPLPAK_B200_RG_WARPS=$w SPLPAK_B200_BASIS=uniform timeout 300 python scripts/eval_ab.py 1e9 3 3 2>&1 | grep regroup | tee -a gpurun_out/r3a_eval_ab_warps.log
done
