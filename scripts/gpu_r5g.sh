#!/bin/bash
# cfg4 assembly: launch list + full captures of the 4-D moment kernel and the cell transform
OUT=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spl_(moments4|cell_transform4|classify|perm|scan|items|wmax|hist)" -c 40 --csv --log-file $OUT/r5g_launches.csv python scripts/cfg4_fit_once.py 1e7 1 > $OUT/r5g_ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spl_moments4 -c 1 -f -o $OUT/r5g_moments4 python scripts/cfg4_fit_once.py 1e7 1 > $OUT/r5g_ncu_m4.log 2>&1
echo "ncu moments4 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spl_cell_transform4 -c 1 -f -o $OUT/r5g_transform4 python scripts/cfg4_fit_once.py 1e7 1 > $OUT/r5g_ncu_t4.log 2>&1
echo "ncu transform4 rc=$?"
