#!/bin/bash
# launch list of the bench command of the last build (two-level partition in place of spl_perm_kernel)
set -u
OUT=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > $OUT/r5y_plain.json 2> $OUT/r5y_plain.err || { echo "plain run failed"; tail -5 $OUT/r5y_plain.err; exit 1; }
NL=$(python -c "import json;print(json.load(open('$OUT/r5y_plain.json'))['gpu_launches']//2)")
echo "launches per step: $NL"
SKIP=$((NL*3))
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:spl_ -s $SKIP -c $((NL*2)) --csv \
    --log-file $OUT/r5y_launches.csv $CMD > $OUT/r5y_ncu_launch.log 2>&1
echo "launch list rc=$?"
