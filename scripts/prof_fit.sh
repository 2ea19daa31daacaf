#!/bin/bash
# short fit-only run for ncu: 2e7 points on the cfg3 grid.  Launch list of one fit + full captures.
OUT=gpurun_out
TAG=${1:-r01}
CMD="python scripts/gpu_time.py 2e7 1e6"
$CMD > $OUT/prof_fit_plain.log 2>&1 || { tail -5 $OUT/prof_fit_plain.log; exit 1; }
head -5 $OUT/prof_fit_plain.log
# one fit = 9 + 3*216 launches (skip the first two fits)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:spl_ -s 1320 -c 660 --csv --log-file $OUT/launches_fit_${TAG}.csv $CMD > $OUT/ncu_launch_fit.log 2>&1
echo "launch list rc=$?"
for K in accumulate classify perm; do
ncu --set full --clock-control none --import-source on -k regex:spl_${K} -s 1 -c 1 -f -o $OUT/prof_${TAG}_${K} $CMD > $OUT/ncu_${TAG}_${K}.log 2>&1
echo "ncu ${K} rc=$?"
done
