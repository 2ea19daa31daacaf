#!/bin/bash
# ncu evidence for the bench command (round 2): launch list of the timed steps + full captures of the dominant kernels.
set -u
OUT=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > $OUT/prof_plain.json 2> $OUT/prof_plain.err || { echo "plain run failed"; tail -5 $OUT/prof_plain.err; exit 1; }
NL=$(python -c "import json;print(json.load(open('$OUT/prof_plain.json'))['gpu_launches']//2)")
echo "launches per step: $NL"
SKIP=$((NL*3))
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:spl_ -s $SKIP -c $((NL*2)) --csv \
    --log-file $OUT/launches_r02.csv $CMD > $OUT/ncu_launch.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:spl_eval_regroup -s 3 -c 1 -f -o $OUT/prof_r02_eval \
    $CMD > $OUT/ncu_r02_eval.log 2>&1
echo "full capture eval rc=$?"
ncu --set full --clock-control none --import-source on -k regex:spl_moments -s 3 -c 1 -f -o $OUT/prof_r02_accumulate \
    $CMD > $OUT/ncu_r02_accumulate.log 2>&1
echo "full capture accumulate rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"spl_factor_(dataflow|persistent)" -s 3 -c 1 -f -o $OUT/prof_r02_factor \
    $CMD > $OUT/ncu_r02_factor.log 2>&1
echo "full capture factor rc=$?"
ls -la $OUT | grep r02
