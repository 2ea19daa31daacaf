#!/bin/bash
# short fit-only run for ncu: 2e7 points on the cfg3 grid
OUT=gpurun_out
CMD="python scripts/gpu_time.py 2e7 1e6"
$CMD > $OUT/prof_fit_plain.log 2>&1 || { tail -5 $OUT/prof_fit_plain.log; exit 1; }
head -5 $OUT/prof_fit_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:spl_ -s 680 -c 680 --csv --log-file $OUT/launches_fit.csv $CMD > $OUT/ncu_launch_fit.log 2>&1
echo "launch list rc=$?"
for K in accumulate panel syrk backsolve; do
ncu --set full --clock-control none --import-source on -k regex:spl_${K} -s 120 -c 1 -f -o $OUT/prof3_${K} $CMD > $OUT/ncu3_${K}.log 2>&1
echo "ncu ${K} rc=$?"
done
