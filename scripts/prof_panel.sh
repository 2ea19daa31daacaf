#!/bin/bash
OUT=gpurun_out
CMD="python scripts/gpu_time.py 2e6 1e6"
$CMD > $OUT/prof_panel_plain.log 2>&1 || { tail -5 $OUT/prof_panel_plain.log; exit 1; }
ncu --set full --sampling-interval min --clock-control none --import-source on -k regex:spl_panel -s 120 -c 2 -f -o $OUT/prof6_panel $CMD > $OUT/ncu6_panel.log 2>&1
echo "ncu rc=$?"
