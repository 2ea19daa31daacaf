#!/bin/bash
# ncu --set full of the evaluation kernel: launch 1 = uniform-random queries, launch 2 = raster order
OUT=gpurun_out
NQ=${1:-1e8}
TAG=${2:-eval}
CMD="python scripts/eval_time.py $NQ 1"
$CMD > $OUT/prof_${TAG}_plain.log 2>&1 || { tail -5 $OUT/prof_${TAG}_plain.log; exit 1; }
cat $OUT/prof_${TAG}_plain.log
ncu --set full --clock-control none --import-source on -k regex:spl_eval -c 2 -f -o $OUT/prof_${TAG} $CMD > $OUT/ncu_${TAG}.log 2>&1
echo "ncu rc=$?"
