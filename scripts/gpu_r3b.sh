#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_regroup.py tests/test_gpu_eval.py -x -q -m gpu > gpurun_out/r3b_tests_eval.log 2>&1; echo "eval tests rc=$?"
tail -5 gpurun_out/r3b_tests_eval.log
for w in 16 24; do
  export SPLPAK_B200_RG_WARPS=$w
  ncu --set full --clock-control none --import-source on -k regex:spl_eval_regroup -c 1 -f -o gpurun_out/prof_r3b_eval_w$w python scripts/eval_time.py 1e8 1 > gpurun_out/ncu_r3b_eval_w$w.log 2>&1
  echo "ncu w$w rc=$?"
done
unset SPLPAK_B200_RG_WARPS
ncu --set full --clock-control none --import-source on -k regex:spl_eval_kernel -c 2 -f -o gpurun_out/prof_r3b_eval_plain python scripts/eval_time.py 1e8 1 > gpurun_out/ncu_r3b_eval_plain.log 2>&1
echo "ncu plain rc=$?"
