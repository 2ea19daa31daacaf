#!/bin/bash
# short eval-only run for ncu (2e8 random queries, cfg3 grid)
OUT=gpurun_out
CMD="python scripts/gpu_time.py 2e6 2e8"
$CMD > $OUT/prof_eval_plain.log 2>&1 || { tail -5 $OUT/prof_eval_plain.log; exit 1; }
tail -3 $OUT/prof_eval_plain.log
ncu --set full --clock-control none --import-source on -k regex:spl_eval -s 2 -c 1 -f -o $OUT/prof_eval2 $CMD > $OUT/ncu_eval2.log 2>&1
echo "ncu rc=$?"
