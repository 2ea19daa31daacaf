#!/bin/bash
# ncu evidence for the bench command: launch list (per-launch device time of our kernels in the timed steps)
# + one --set full capture each of the two dominant kernels at bench size.  Run under gpurun; outputs land in
# gpurun_out/; scripts/profile_digest.py turns them into profiles/*.md and profiles/r01_traffic.json.
set -u
OUT=gpurun_out
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > $OUT/prof_plain.json 2> $OUT/prof_plain.err || { echo "plain run failed"; tail -5 $OUT/prof_plain.err; exit 1; }
NL=$(python -c "import json;print(json.load(open('$OUT/prof_plain.json'))['gpu_launches']//2)")
echo "launches per step: $NL"
SKIP=$((NL*3))
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:spl_ -s $SKIP -c $((NL*2)) --csv \
    --log-file $OUT/launches_${TAG}.csv $CMD > $OUT/ncu_launch.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:spl_eval -s 3 -c 1 -f -o $OUT/prof_${TAG}_eval \
    $CMD > $OUT/ncu_${TAG}_eval.log 2>&1
echo "full capture eval rc=$?"
ncu --set full --clock-control none --import-source on -k regex:spl_moments -s 3 -c 1 -f -o $OUT/prof_${TAG}_accumulate \
    $CMD > $OUT/ncu_${TAG}_accumulate.log 2>&1
echo "full capture accumulate rc=$?"
# the panel code runs inside the persistent factor kernel by default; SPLPAK_B200_SOLVER=graph launches it per panel
SPLPAK_B200_SOLVER=graph ncu --set full --clock-control none --import-source on -k regex:spl_panel -s 700 -c 1 -f -o $OUT/prof_${TAG}_panel \
    $CMD > $OUT/ncu_${TAG}_panel.log 2>&1
echo "full capture panel rc=$?"
ls -la $OUT | tail -8
