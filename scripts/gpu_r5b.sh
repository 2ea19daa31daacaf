#!/bin/bash
# cfg4 factor loop (K-blocked): launch list and one full capture of the K-blocked update (rest part)
OUT=gpurun_out
CMD="python scripts/cfg4_fit_once.py 1e6 2"
$CMD > $OUT/r5b_plain.log 2>&1 || { tail -5 $OUT/r5b_plain.log; exit 1; }
cat $OUT/r5b_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spl_(panel|syrk)" -c 800 --csv --log-file $OUT/r5b_launches.csv python scripts/cfg4_fit_once.py 1e6 1 > $OUT/r5b_ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spl_syrk_kblock -s 81 -c 1 -f -o $OUT/r5b_kblock python scripts/cfg4_fit_once.py 1e6 1 > $OUT/r5b_ncu_kblock.log 2>&1
echo "ncu kblock rc=$?"
