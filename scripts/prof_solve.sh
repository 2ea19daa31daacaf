#!/bin/bash
OUT=gpurun_out
CMD="python scripts/gpu_time.py 2e6 1e6"
$CMD > $OUT/prof_solve_plain.log 2>&1 || { tail -5 $OUT/prof_solve_plain.log; exit 1; }
head -2 $OUT/prof_solve_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:spl_ -s 680 -c 680 --csv --log-file $OUT/launches_solve.csv $CMD > $OUT/ncu_launch_solve.log 2>&1
for K in panel syrk; do
ncu --set full --clock-control none --import-source on -k regex:spl_${K} -s 120 -c 1 -f -o $OUT/prof5_${K} $CMD > $OUT/ncu5_${K}.log 2>&1
echo "ncu ${K} rc=$?"
done
