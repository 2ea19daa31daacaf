#!/bin/bash
# ncu --set full of one kernel (regex $1) of the cfg3 fit at ${NPTS:-1e8} points -> gpurun_out/prof_$2.ncu-rep
OUT=gpurun_out
K=$1
TAG=${2:-k}
CMD="python scripts/gpu_time.py ${NPTS:-1e8} 1e6"
$CMD > $OUT/prof_${TAG}_plain.log 2>&1 || { tail -5 $OUT/prof_${TAG}_plain.log; exit 1; }
head -3 $OUT/prof_${TAG}_plain.log
ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o $OUT/prof_${TAG} $CMD > $OUT/ncu_${TAG}.log 2>&1
echo "ncu rc=$?"
