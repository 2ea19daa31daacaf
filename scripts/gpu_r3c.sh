#!/bin/bash
mkdir -p gpurun_out
for w in 8 12 16 24; do
  echo "== uniform, $w warps"
  SPLPAK_B200_RG_WARPS=$w timeout 300 python scripts/eval_ab.py 1e9 3 3 2>&1 | grep "random regroup" | tee -a gpurun_out/r3c_eval_ab_warps.log
done
timeout 300 python scripts/eval_ab.py 1e9 3 1,4 2>&1 | tee gpurun_out/r3c_eval_ab_14.log
SPLPAK_B200_BASIS=exact timeout 300 python scripts/eval_ab.py 1e9 3 1 2>&1 | tee gpurun_out/r3c_eval_ab_1exact.log
