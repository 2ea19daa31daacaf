#!/bin/bash
# (1) launch list of the bench command of the final build; (2) full capture of the bulk (part 1) K-blocked update at cfg4
set -u
OUT=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > $OUT/r5n_plain.json 2> $OUT/r5n_plain.err || { echo "plain run failed"; tail -5 $OUT/r5n_plain.err; exit 1; }
NL=$(python -c "import json;print(json.load(open('$OUT/r5n_plain.json'))['gpu_launches']//2)")
echo "launches per step: $NL"
SKIP=$((NL*3))
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:spl_ -s $SKIP -c $((NL*2)) --csv \
    --log-file $OUT/r5n_launches.csv $CMD > $OUT/r5n_ncu_launch.log 2>&1
echo "launch list rc=$?"
# cfg4: the 20th launch of spl_syrk_kblock_kernel is a bulk part (parts alternate: rest, column part, rest, ...)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spl_syrk_kblock -s 20 -c 2 -f -o $OUT/r5n_kblock python scripts/cfg4_fit_once.py 1e6 1 > $OUT/r5n_ncu_kblock.log 2>&1
echo "ncu kblock rc=$?"
