#!/bin/bash
# paired back-substitution: solver checks in every driver, solver / scale / refinement tests, bench
mkdir -p gpurun_out
for mode in dataflow graph; do
  if [ $mode = dataflow ]; then unset SPLPAK_B200_SOLVER; else export SPLPAK_B200_SOLVER=$mode; fi
  echo "== solver_check $mode"; timeout 300 python scripts/solver_check.py 2>&1 | tail -9
done
unset SPLPAK_B200_SOLVER
timeout 900 python -m pytest tests/test_gpu_fit.py tests/test_gpu_scale.py tests/test_gpu_ortho.py -x -q -m gpu > gpurun_out/r4n_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r4n_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r4n_bench.json 2> gpurun_out/r4n_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r4n_bench.json') if l.startswith('{')][-1]); s=d['stages_ms']; print('fit %.2f eval %.2f' % (d['fit_ms'], d['eval_ms']), {k: round(v,2) for k,v in s.items()}, 'chk', d['checksum'])"
timeout 600 python scripts/config_times.py > gpurun_out/r4n_configs.md 2>&1; tail -4 gpurun_out/r4n_configs.md
