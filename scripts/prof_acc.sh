#!/bin/bash
OUT=gpurun_out
CMD="python scripts/gpu_time.py 2e7 1e6"
$CMD > $OUT/prof_acc_plain.log 2>&1 || { tail -5 $OUT/prof_acc_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:spl_accumulate -s 1 -c 1 -f -o $OUT/prof4_accumulate $CMD > $OUT/ncu4_accumulate.log 2>&1
echo "ncu rc=$?"
