#!/bin/bash
# ncu --set full of the moment kernel on 2e7 cfg3 points
OUT=gpurun_out
TAG=${1:-mom}
CMD="python scripts/gpu_time.py ${NPTS:-2e7} 1e6"
$CMD > $OUT/prof_${TAG}_plain.log 2>&1 || { tail -5 $OUT/prof_${TAG}_plain.log; exit 1; }
head -3 $OUT/prof_${TAG}_plain.log
ncu --set full --clock-control none --import-source on -k regex:spl_moments -s 1 -c 1 -f -o $OUT/prof_${TAG} $CMD > $OUT/ncu_${TAG}.log 2>&1
echo "ncu rc=$?"
