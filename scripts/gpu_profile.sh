#!/bin/bash
# ncu evidence for one bench command: launch list (per-launch device time of our kernels) + one
# --set full capture of the top kernels.  Run under gpurun; outputs land in gpurun_out/.
set -u
OUT=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > $OUT/prof_plain.json 2> $OUT/prof_plain.err || { echo "plain run failed"; tail -5 $OUT/prof_plain.err; exit 1; }
# launches of our kernels in the 2 timed steps (skip the 3 warm-up steps' launches)
NL=$(python -c "import json;print(json.load(open('$OUT/prof_plain.json'))['gpu_launches']//2)")
SKIP=$((NL*3))
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:spl_ -s $SKIP -c $((NL*2)) --csv \
    --log-file $OUT/launches.csv $CMD > $OUT/ncu_launch.log 2>&1
echo "launch list rc=$?"
for K in eval accumulate scatter panel syrk backsolve classify; do
  ncu --set full --clock-control none --import-source on -k regex:spl_${K} -s 4 -c 1 -f -o $OUT/prof_${K} \
      $CMD > $OUT/ncu_${K}.log 2>&1
  echo "full capture ${K} rc=$?"
done
ls -la $OUT
